"""Small GPU driver for profiling: N forwards of the BASELINE model at batch B (direct launches unless ISHARA_GRAPH=1).
usage: python tools/fwd_once.py [B] [N]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("ISHARA_GRAPH", "0")
import numpy as np

import ishara_b200 as ib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m = ib.get_model()
rng = np.random.default_rng(0)
x = rng.standard_normal((B, 384, 276), dtype=np.float32)
for i in range(N):
    t0 = time.time()
    y = m(x)
    print("forward", i, y.shape, float(np.abs(y).max()), f"{time.time() - t0:.3f}s")
