"""sylt2d_b200 — B200-native (sm_100a, hand-written CUDA) implementation of sylt-2d's World::step hot path.

The product is the C-ABI shared library libsylt2d_b200.so (include/sylt2d_b200.h); this package holds its CUDA
sources (csrc/), the Python mirror of the reference's public API used by the tests (world.py, batch.py) and the
scene library (scenes.py).  Importing the package never touches the GPU; `lib()` raises if the library is missing.
"""
from . import scenes  # noqa: F401
from ._ffi import SyltError, lib  # noqa: F401
from .batch import Batch, shard_range  # noqa: F401
from .world import Arbiter, Body, Joint, Vec2, World, WorldContext, collide_pairs, device_sincos  # noqa: F401
