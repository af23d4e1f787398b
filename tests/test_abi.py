"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/sylt2d_b200.h declares; record layouts match the header; host-side logic (scene generators, sharding,
joint colouring inputs) behaves.  No compute entry point is called here (no GPU in this container)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from sylt2d_b200 import _ffi
    return _ffi


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "sylt2d_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sylt_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 50
    L = built.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(built.SYMBOLS) == declared


def test_record_layouts_match_header(built):
    from sylt2d_b200.scenes import BODY_DESC, JOINT_DESC
    sizes = {"SyltBodyDesc": 64, "SyltJointDesc": 32, "SyltBodyState": 36, "SyltContact": 68, "SyltArbiterInfo": 40, "SyltJointState": 80,
             "SyltStepStats": 32, "SyltBatchStats": 40}
    src = "#include <stdio.h>\n#include \"sylt2d_b200.h\"\nint main(){" + "".join(
        f'printf("{k} %zu\\n", sizeof({k}));' for k in sizes) + "return 0;}"
    import subprocess
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout
    got = dict(line.split() for line in out.strip().splitlines())
    assert {k: int(v) for k, v in got.items()} == sizes
    assert BODY_DESC.itemsize == 64 and JOINT_DESC.itemsize == 32
    assert built.BODY_STATE.itemsize == 36 and built.CONTACT.itemsize == 68 and built.ARBITER.itemsize == 40


def test_no_cpu_fallback_without_device(built):
    L = built.lib()
    if L.sylt_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert L.sylt_world_create(0.0, -10.0, 10, 0, C.byref(h)) == built.ERR_CUDA  # fails loudly, no silent CPU path
    assert h.value is None


def test_product_does_not_link_or_import_the_oracle():
    import subprocess
    so = os.path.join(ROOT, "sylt2d_b200", "libsylt2d_b200.so")
    out = subprocess.run(["nm", "-D", so], capture_output=True, text=True).stdout
    assert "orc_" not in out
    for f in ("world.py", "batch.py", "_ffi.py", "scenes.py", "__init__.py", "build.py"):
        src = open(os.path.join(ROOT, "sylt2d_b200", f)).read()
        assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M) and "liborc" not in src and "orc." not in src, f


def test_scene_generators_match_the_samples():
    from sylt2d_b200 import scenes
    p12 = scenes.pyramid(12, 100)  # demo5 literal: 78 boxes + ground (examples/samples/src/main.rs:157-178)
    assert p12.num_bodies == 79 and p12.iterations == 100
    b = p12.bodies
    assert b[0]["mass"] == np.float32(scenes.F32_MAX) and (b[0]["x"], b[0]["y"]) == (0.0, -10.0)
    assert (b[1]["x"], b[1]["y"]) == (-6.0, 0.75) and b[2]["x"] == np.float32(-6.0 + 1.125)
    assert (b[13]["x"], b[13]["y"]) == (np.float32(-6.0 + 0.5625), 2.75)   # first box of row 1
    assert scenes.pyramid(20, 10).num_bodies == 211
    br = scenes.bridge()
    assert br.num_bodies == 17 and br.joints.shape[0] == 16
    assert all(int(j["id2"]) == 1 for j in br.joints)                      # quirk Q-13: every plank pinned to the ground
    assert abs(br.joints[0]["softness"] - 1.0 / (2 * 10 * 0.7 * 4 * np.pi + (1 / 60) * 10 * (4 * np.pi) ** 2)) < 1e-6
    mp = scenes.multi_pendulum()
    assert mp.num_bodies == 16 and mp.joints.shape[0] == 15 and int(mp.joints[1]["id1"]) == 2
    assert np.all(np.diff(scenes.mixed_pile(300).bodies["id"].astype(np.int64)) == 1)  # increasing ids (quirk Q-6)
    d = scenes.box_drop(1000)
    assert d.num_bodies == 1003


def test_tile_worlds_and_shard_range():
    from sylt2d_b200 import scenes
    from sylt2d_b200.batch import shard_range
    sc = scenes.bridge()
    bodies, boff, joints, joff = scenes.tile_worlds(sc, 5, seed=1, omega_jitter=1.0)
    assert bodies.shape[0] == 5 * 17 and joints.shape[0] == 5 * 16 and list(boff) == [0, 17, 34, 51, 68, 85]
    w = bodies["angular_velocity"].reshape(5, 17)
    assert np.all(w[:, 1] != 0) and np.all(w[:, 2:] == 0)
    cover = []
    for ws in (1, 2, 3, 4, 8):
        got = [shard_range(10, r, ws) for r in range(ws)]
        assert got[0][0] == 0 and max(h for _, h in got) == 10
        flat = [k for lo, hi in got for k in range(lo, hi)]
        assert flat == list(range(10))
        cover.append(got)


def _build_cpp_example(tmpdir):
    import subprocess
    exe = os.path.join(tmpdir, "pyramid_cpp")
    lib = os.path.join(ROOT, "sylt2d_b200")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "pyramid.cpp"),
                    "-L", lib, "-lsylt2d_b200", f"-Wl,-rpath,{lib}", "-o", exe], check=True)
    return exe


def test_cpp_mirror_compiles_and_links(built, tmp_path):
    import subprocess
    exe = _build_cpp_example(str(tmp_path))
    L = built.lib()
    if L.sylt_device_count() == 0:
        r = subprocess.run([exe, "3"], capture_output=True, text=True)
        assert r.returncode == 2 and "error 6" in r.stderr  # SYLT_ERR_CUDA surfaced as sylt2d::Error, no CPU fallback
