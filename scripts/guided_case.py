#!/usr/bin/env python3
"""2^22 eval-guided playouts from the empty board: ncu target."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantik-core-py_b200"))
import quantik_b200 as Q  # noqa: E402

e = np.zeros((1, 8), np.uint16)
Q.guided_playouts(e, 1 << 16)
r = Q.guided_playouts(e, 1 << 22)
print(int(r["wins_p0"][0]), int(r["wins_p1"][0]))
