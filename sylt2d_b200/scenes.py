"""Scene library: the reference's sample scenes restated as data.

Each generator returns a `Scene` whose `bodies` / `joints` are numpy structured arrays laid out exactly as
`SyltBodyDesc` / `SyltJointDesc` of include/sylt2d_b200.h, so that a scene can be handed to the C ABI in
one bulk call.  Geometry follows /root/reference/examples/samples/src/main.rs (demo5 pyramid :157-178,
demo7 bridge :206-239, demo9 multi-pendulum :300-336, demo1..demo10 :66-391) with f32 arithmetic in the same
order as the Rust source.  Bodies are created in increasing-id order (reference quirk Q-6).
"""
from dataclasses import dataclass, field
import numpy as np

F32_MAX = float(np.finfo(np.float32).max)

BODY_DESC = np.dtype([("id", "<u8"), ("shape", "<i4"), ("vert_offset", "<i4"), ("vert_count", "<i4"), ("group", "<i4"),
                      ("width_x", "<f4"), ("width_y", "<f4"), ("mass", "<f4"), ("friction", "<f4"),
                      ("x", "<f4"), ("y", "<f4"), ("rotation", "<f4"), ("vx", "<f4"), ("vy", "<f4"), ("angular_velocity", "<f4")])
JOINT_DESC = np.dtype([("id1", "<u8"), ("id2", "<u8"), ("anchor_x", "<f4"), ("anchor_y", "<f4"),
                       ("softness", "<f4"), ("bias_factor", "<f4")])
assert BODY_DESC.itemsize == 64 and JOINT_DESC.itemsize == 32

f32 = np.float32


@dataclass
class Scene:
    name: str
    bodies: np.ndarray
    joints: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=JOINT_DESC))
    verts: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), dtype=np.float32))
    gravity: tuple = (0.0, -10.0)       # examples/samples/src/main.rs:39
    iterations: int = 10
    dt: float = float(f32(1.0) / f32(60.0))  # :45
    ctx: tuple = (True, True, True)     # (accumulate_impulse, warm_starting, position_correction)

    @property
    def num_bodies(self):
        return int(self.bodies.shape[0])


class _Builder:
    def __init__(self, first_id=1):
        self.rows = []
        self.joints = []
        self.verts = []
        self.nv = 0
        self.next_id = first_id

    def box(self, w, h, mass, friction=0.0, pos=(0.0, 0.0), rot=0.0, vel=(0.0, 0.0), omega=0.0):
        i = self.next_id
        self.next_id += 1
        self.rows.append((i, 0, 0, 0, 0, w, h, mass, friction, pos[0], pos[1], rot, vel[0], vel[1], omega))
        return i

    def polygon(self, verts, mass, friction=0.0, pos=(0.0, 0.0), rot=0.0, vel=(0.0, 0.0), omega=0.0):
        v = np.asarray(verts, dtype=np.float32).reshape(-1, 2)
        i = self.next_id
        self.next_id += 1
        self.rows.append((i, 1, self.nv, v.shape[0], 0, 0.0, 0.0, mass, friction, pos[0], pos[1], rot, vel[0], vel[1], omega))
        self.verts.append(v)
        self.nv += v.shape[0]
        return i

    def joint(self, id1, id2, anchor, softness=0.0, bias_factor=0.2):
        self.joints.append((id1, id2, anchor[0], anchor[1], softness, bias_factor))

    def scene(self, name, **kw):
        b = np.array(self.rows, dtype=BODY_DESC) if self.rows else np.zeros(0, dtype=BODY_DESC)
        j = np.array(self.joints, dtype=JOINT_DESC) if self.joints else np.zeros(0, dtype=JOINT_DESC)
        v = np.concatenate(self.verts).astype(np.float32) if self.verts else np.zeros((0, 2), dtype=np.float32)
        return Scene(name=name, bodies=b, joints=j, verts=v, **kw)


def _ground(b, w=100.0, h=20.0, friction=0.2):
    # Body::new((100,20), f32::MAX) at (0, -0.5*width.y)   (examples/samples/src/main.rs:159-162)
    return b.box(w, h, F32_MAX, friction, (0.0, float(f32(-0.5) * f32(h))))


def pyramid(rows=20, iterations=10, x_jitter=None, rng=None):
    """demo5 (examples/samples/src/main.rs:157-178) with `12 -> rows`; config 1/5 use rows=20, I=10."""
    b = _Builder()
    _ground(b)
    x = [f32(-6.0), f32(0.75)]
    k = 0
    for i in range(rows):
        y = [x[0], x[1]]
        for _ in range(i, rows):
            px = y[0]
            if x_jitter is not None:
                px = f32(px + f32(rng.uniform(-x_jitter, x_jitter)))
            b.box(1.0, 1.0, 10.0, 0.2, (float(px), float(y[1])))
            y[0] = f32(y[0] + f32(1.125))
            k += 1
        x[0] = f32(x[0] + f32(0.5625))
        x[1] = f32(x[1] + f32(2.0))
    return b.scene(f"pyramid{rows}", iterations=iterations)


def _soft_params(mass, frequency_hz, damping_ratio, time_step):
    # examples/samples/src/main.rs:215-221 / :311-318, f32 left-to-right
    mass, frequency_hz, damping_ratio, time_step = f32(mass), f32(frequency_hz), f32(damping_ratio), f32(time_step)
    omega = f32(f32(f32(2.0) * f32(np.pi)) * frequency_hz)
    d = f32(f32(f32(f32(2.0) * mass) * damping_ratio) * omega)
    k = f32(f32(mass * omega) * omega)
    softness = f32(f32(1.0) / f32(d + f32(time_step * k)))
    bias_factor = f32(f32(time_step * k) / f32(d + f32(time_step * k)))
    return float(softness), float(bias_factor)


def bridge(iterations=10, omega0=0.0):
    """demo7 (examples/samples/src/main.rs:206-239): 16 planks, each pinned to the GROUND (quirk Q-13)."""
    b = _Builder()
    g = _ground(b)
    softness, bias = _soft_params(10.0, 2.0, 0.7, f32(1.0) / f32(60.0))
    for i in range(16):
        fi = f32(i)
        p = b.box(1.0, 0.25, 10.0, 0.2, (float(f32(-8.5) + f32(1.25) * fi), 5.0), omega=omega0 if i == 0 else 0.0)
        b.joint(p, g, (float(f32(-9.125) + f32(1.25) * fi), 5.0), softness, bias)
    return b.scene("bridge", iterations=iterations)


def multi_pendulum(iterations=10, omega0=0.0):
    """demo9 (examples/samples/src/main.rs:300-336): chain of 15 links."""
    b = _Builder()
    prev = _ground(b)
    softness, bias = _soft_params(10.0, 4.0, 0.7, f32(1.0) / f32(60.0))
    for i in range(15):
        fi = f32(i)
        p = b.box(0.75, 0.25, 10.0, 0.2, (float(f32(0.5) + fi), 12.0), omega=omega0 if i == 0 else 0.0)
        b.joint(prev, p, (float(fi), 12.0), softness, bias)
        prev = p
    return b.scene("multi_pendulum", iterations=iterations)


HEXAGON = [(0.0, 1.0), (-0.87, 0.5), (-0.87, -0.5), (0.0, -1.0), (0.87, -0.5), (0.87, 0.5)]   # main.rs:77-84
PENTAGON = [(0.0, 1.0), (-0.95, 0.31), (-0.59, -0.81), (0.59, -0.81), (0.95, 0.31)]           # main.rs:86-92


def demo1(iterations=10):
    """Single shapes falling (main.rs:66-101): ground, a box, a pentagon and a hexagon."""
    b = _Builder()
    _ground(b, friction=0.0)
    b.box(1.0, 1.0, 200.0, 0.0, (0.0, 3.0))
    b.polygon(PENTAGON, 2.0, 100.0, (0.0, 5.0))
    b.polygon(HEXAGON, 2.0, 0.0, (5.0, 4.0), rot=45.0)
    return b.scene("demo1", iterations=iterations)


def demo2(iterations=10):
    """Simple pendulum (main.rs:103-118)."""
    b = _Builder()
    g = _ground(b)
    p = b.box(1.0, 1.0, 100.0, 0.2, (9.0, 11.0))
    b.joint(g, p, (0.0, 11.0))
    return b.scene("demo2", iterations=iterations)


def demo3(iterations=10):
    """Varying friction (main.rs:120-138)."""
    b = _Builder()
    _ground(b, friction=0.0)
    b.box(13.0, 0.25, F32_MAX, 0.0, (-2.0, 11.0), rot=-0.25)
    for i, fr in enumerate([0.75, 0.5, 0.35, 0.1, 0.0]):
        b.box(0.5, 0.5, 25.0, fr, (float(f32(-7.5) + f32(2.0) * f32(i)), 14.0))
    return b.scene("demo3", iterations=iterations)


def demo4(iterations=10, seed=0):
    """Vertical stack (main.rs:140-154); the sample's random x offset is seeded here."""
    rng = np.random.default_rng(seed)
    b = _Builder()
    _ground(b)
    for i in range(10):
        b.box(1.0, 1.0, 1.0, 0.2, (float(f32(rng.random()) * f32(0.2) - f32(0.1)), float(f32(0.51) + f32(1.05) * f32(i))))
    return b.scene("demo4", iterations=iterations)


def demo6(iterations=10):
    """Teeter (main.rs:180-204)."""
    b = _Builder()
    g = _ground(b, friction=0.0)
    t = b.box(12.0, 0.25, 10.0, 0.0, (0.0, 3.0))
    b.box(0.5, 0.5, 2.0, 0.0, (-5.0, 5.0))
    b.box(0.5, 0.5, 2.0, 0.0, (-5.5, 5.0))
    b.box(1.0, 1.0, 55.0, 0.0, (5.5, 15.0))
    b.joint(g, t, (0.0, 3.0))
    return b.scene("demo6", iterations=iterations)


def demo8(iterations=10):
    """Dominos (main.rs:242-297)."""
    b = _Builder()
    b1 = _ground(b, friction=0.0)
    b.box(12.0, 0.5, F32_MAX, 0.0, (-1.5, 10.0))
    for i in range(10):
        b.box(0.2, 2.0, 10.0, 0.1, (float(f32(-6.0) + f32(1.0) * f32(i)), 11.125))
    b.box(14.0, 0.5, F32_MAX, 0.0, (1.0, 6.0), rot=0.3)
    b2 = b.box(0.5, 3.0, F32_MAX, 0.0, (-7.0, 4.0))
    b3 = b.box(12.0, 0.25, 10.0, 0.0, (-0.9, 1.0))
    b.joint(b1, b3, (-2.0, 3.0))
    b4 = b.box(0.5, 0.5, 16.0, 0.2, (-10.0, 15.0))
    b.joint(b2, b4, (-7.0, 15.0))
    b5 = b.box(2.0, 2.0, 10.0, 0.1, (6.0, 2.5))
    b.joint(b1, b5, (6.0, 2.6))
    b6 = b.box(2.0, 0.2, 10.0, 0.0, (6.0, 3.6))
    b.joint(b5, b6, (7.0, 3.5))
    return b.scene("demo8", iterations=iterations)


def _scaled(verts, factor):
    # ConvexPolygon::scale (src/body.rs:84-93) — about the (shifted-index) centroid, f32
    v = np.asarray(verts, dtype=np.float32)
    n = v.shape[0]
    cx = cy = area = f32(0.0)
    for i in range(n):
        p1, p2 = v[(i + 1) % n], v[(i + 2) % n]
        cr = f32(p1[0] * p2[1]) - f32(p1[1] * p2[0])
        area = f32(area + cr)
        cx = f32(cx + f32(f32(p1[0] + p2[0]) * cr))
        cy = f32(cy + f32(f32(p1[1] + p2[1]) * cr))
    area = f32(area / f32(2.0))
    cx = f32(cx / f32(f32(6.0) * area))
    cy = f32(cy / f32(f32(6.0) * area))
    c = np.array([cx, cy], dtype=np.float32)
    return ((v - c) * f32(factor) + c).astype(np.float32)


def demo10(iterations=10):
    """Pawn and pendulum (main.rs:339-391): polygons joined by joints."""
    head = _scaled([(0.0, 0.4), (-0.4, 0.0), (-0.2, -0.4), (0.2, -0.4), (0.4, 0.0)], 2.0)
    trunk = _scaled([(-0.8, 0.8), (-0.6, 0.0), (-0.3, -0.8), (0.3, -0.8), (0.6, 0.0), (0.8, 0.8)], 2.0)
    b = _Builder()
    g = b.box(1000.0, 20.0, F32_MAX, 0.0, (0.0, -10.0))
    wt = trunk.max(0) - trunk.min(0)
    wh = head.max(0) - head.min(0)
    pent = b.polygon(PENTAGON, 55.0, 0.2, (-9.0, 8.0))
    hy = float(f32(f32(4.0) + f32(f32(0.5) * wt[1])) + f32(f32(0.5) * wh[1]))
    hd = b.polygon(head, 10.0, 0.0, (5.0, hy))
    tr = b.polygon(trunk, 10.0, 0.0, (5.0, 4.0))
    b.joint(hd, tr, (5.0, 3.0))
    b.joint(g, pent, (0.0, 11.0))
    return b.scene("demo10", iterations=iterations)


def regular_ngon(n, radius, phase=0.0):
    """CCW regular n-gon (quirk Q-8: input polygons must already be counter-clockwise)."""
    a = phase + 2.0 * np.pi * np.arange(n) / n
    return np.stack([radius * np.cos(a), radius * np.sin(a)], axis=1).astype(np.float32)


def _container(b, width, height, thick=2.0, friction=0.4):
    # 3-wall static container: floor + two side walls (static: mass == f32::MAX, src/body.rs:208-216)
    b.box(width + 2 * thick, thick, F32_MAX, friction, (0.0, -0.5 * thick))
    b.box(thick, height, F32_MAX, friction, (-0.5 * width - 0.5 * thick, 0.5 * height))
    b.box(thick, height, F32_MAX, friction, (0.5 * width + 0.5 * thick, 0.5 * height))


def mixed_pile(n=10000, seed=1234, iterations=10, across=120):
    """Config 2: seeded mixed box + polygon pile above a 3-wall container."""
    rng = np.random.default_rng(seed)
    pitch = 2.0
    rows = (n + across - 1) // across
    b = _Builder()
    _container(b, across * pitch + 2.0, rows * pitch * 0.6 + 10.0)
    x0 = -0.5 * (across - 1) * pitch
    for k in range(n):
        cx = x0 + (k % across) * pitch + rng.uniform(-0.2, 0.2)
        cy = 1.5 + (k // across) * pitch + rng.uniform(-0.2, 0.2)
        rot = rng.uniform(-np.pi, np.pi)
        fr = rng.uniform(0.1, 0.6)
        if rng.random() < 0.5:
            w, h = rng.uniform(0.5, 1.5), rng.uniform(0.5, 1.5)
            b.box(float(f32(w)), float(f32(h)), float(f32(10.0 * w * h)), float(f32(fr)), (cx, cy), rot)
        else:
            nv = int(rng.integers(3, 9))
            r = rng.uniform(0.4, 0.8)
            area = 0.5 * nv * r * r * np.sin(2 * np.pi / nv)
            b.polygon(regular_ngon(nv, r), float(f32(10.0 * area)), float(f32(fr)), (cx, cy), rot)
    return b.scene(f"mixed_pile{n}", iterations=iterations)


def box_drop(n=100000, seed=42, iterations=10, across=400, pitch=1.25):
    """Config 3: n boxes (1x1 ±10 %), jittered lattice `across` wide, above floor + 2 walls."""
    rng = np.random.default_rng(seed)
    rows = (n + across - 1) // across
    b = _Builder()
    _container(b, across * pitch + 2.0, rows * pitch + 20.0, friction=0.2)
    k = np.arange(n)
    x0 = -0.5 * (across - 1) * pitch
    cx = (x0 + (k % across) * pitch + rng.uniform(-0.05, 0.05, n)).astype(np.float32)
    cy = (1.0 + (k // across) * pitch + rng.uniform(-0.05, 0.05, n)).astype(np.float32)
    w = rng.uniform(0.9, 1.1, n).astype(np.float32)
    h = rng.uniform(0.9, 1.1, n).astype(np.float32)
    first = b.next_id
    sc = b.scene("tmp", iterations=iterations)
    rows_arr = np.zeros(n, dtype=BODY_DESC)
    rows_arr["id"] = first + k
    rows_arr["width_x"], rows_arr["width_y"] = w, h
    rows_arr["mass"], rows_arr["friction"] = 10.0, 0.2
    rows_arr["x"], rows_arr["y"] = cx, cy
    sc.bodies = np.concatenate([sc.bodies, rows_arr])
    sc.name = f"box_drop{n}"
    return sc


def tile_worlds(scene, n_worlds, seed=0, x_jitter=0.0, omega_jitter=0.0):
    """Replicate one scene into `n_worlds` independent worlds (configs 4 and 5).

    Returns (bodies, body_offsets, joints, joint_offsets, verts): CSR arrays for sylt_batch_create.
    Per-world perturbation: x jitter U(-x_jitter, x_jitter) on every dynamic body, and/or an initial angular
    velocity U(-omega_jitter, omega_jitter) on body index 1 — drawn from default_rng(seed + world).
    """
    nb, nj = scene.bodies.shape[0], scene.joints.shape[0]
    bodies = np.tile(scene.bodies, n_worlds)
    joints = np.tile(scene.joints, n_worlds)
    if scene.verts.shape[0]:
        raise ValueError("tile_worlds: polygon scenes are not batched")
    dyn = scene.bodies["mass"] < F32_MAX
    rng = np.random.default_rng(seed)
    if x_jitter:
        jit = rng.uniform(-x_jitter, x_jitter, (n_worlds, nb)).astype(np.float32) * dyn[None, :]
        bodies["x"] = (bodies["x"].reshape(n_worlds, nb) + jit).reshape(-1)
    if omega_jitter and nb > 1:
        om = rng.uniform(-omega_jitter, omega_jitter, n_worlds).astype(np.float32)
        w = bodies["angular_velocity"].reshape(n_worlds, nb).copy()
        w[:, 1] = om
        bodies["angular_velocity"] = w.reshape(-1)
    body_off = (np.arange(n_worlds + 1, dtype=np.int64) * nb).astype(np.int32)
    joint_off = (np.arange(n_worlds + 1, dtype=np.int64) * nj).astype(np.int32)
    return bodies, body_off, joints, joint_off


def group_worlds(scene, n_worlds, seed=0, x_jitter=0.0, omega_jitter=0.0):
    """`n_worlds` independent copies of one scene inside ONE World: copy k gets collision group k (SyltBodyDesc.group), so
    its bodies never meet the other copies', and ids k * stride + id (unique and increasing, as World::add_body needs).
    Unlike tile_worlds / Batch this also takes polygon scenes.  Returns (scene, stride)."""
    nb, nj = scene.bodies.shape[0], scene.joints.shape[0]
    if n_worlds > 65536:
        raise ValueError("at most 65536 collision groups")
    stride = int(scene.bodies["id"].max()) + 1
    bodies = np.tile(scene.bodies, n_worlds)
    joints = np.tile(scene.joints, n_worlds)
    k = np.repeat(np.arange(n_worlds, dtype=np.uint64), nb)
    bodies["id"] = bodies["id"] + k * np.uint64(stride)
    bodies["group"] = k.astype(np.int32)
    if nj:
        kj = np.repeat(np.arange(n_worlds, dtype=np.uint64), nj)
        joints["id1"] = joints["id1"] + kj * np.uint64(stride)
        joints["id2"] = joints["id2"] + kj * np.uint64(stride)
    dyn = scene.bodies["mass"] < F32_MAX
    rng = np.random.default_rng(seed)
    if x_jitter:
        jit = rng.uniform(-x_jitter, x_jitter, (n_worlds, nb)).astype(np.float32) * dyn[None, :]
        bodies["x"] = (bodies["x"].reshape(n_worlds, nb) + jit).reshape(-1)
    if omega_jitter and nb > 1:
        w = bodies["angular_velocity"].reshape(n_worlds, nb).copy()
        w[:, 1] = rng.uniform(-omega_jitter, omega_jitter, n_worlds).astype(np.float32)
        bodies["angular_velocity"] = w.reshape(-1)
    return Scene(f"{scene.name}x{n_worlds}", bodies, joints, scene.verts, scene.gravity, scene.iterations, scene.dt, scene.ctx), stride
