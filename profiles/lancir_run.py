#!/usr/bin/env python
"""One CLancIR 8K->4K RGBA u8 call through the host API (for ncu: -k regex:lancir)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import avir_b200 as ab

rng = np.random.default_rng(1)
src = rng.integers(0, 256, (4320, 7680, 4), dtype=np.uint8)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    r, out = ab.CLancIR().resizeImage(src, 3840, 2160)
    assert r == 2160, r
print("ok", out.shape, int(out.sum()) & 0xffff)
