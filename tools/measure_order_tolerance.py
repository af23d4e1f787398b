#!/usr/bin/env python
"""Measures how much the GPU's colour order differs from the canonical ascending-(id1, id2) order (needs a GPU):
the numbers tests/test_gpu_fullsize.py::ORDER_TOL is derived from (gate = 2 x measured).  Usage:
    python tools/measure_order_tolerance.py > profiles/r04_order_tolerance.json"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_fullsize import measure_order_effects  # noqa: E402

print(json.dumps(measure_order_effects(600), indent=1))
