"""Times ishara_preprocess on device-resident frames (256 sequences, 100..800 raw frames each) with CUDA events."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ishara_b200 import _lib  # noqa: E402
from ishara_b200.preprocess import _flatten_stats  # noqa: E402
from oracle import ishara_preprocess_oracle as PO  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
B, T, F = 256, 384, 276
rng = np.random.default_rng(5)
plens = rng.integers(100, 801, size=B)
offs = np.zeros(B + 1, np.int32)
offs[1:] = np.cumsum(plens)
raw = torch.rand(int(offs[-1]), F, device=dev)
raw[torch.rand(int(offs[-1]), device=dev) < 0.3, :42] = float("nan")
pm, ps = (torch.from_numpy(a).to(dev) for a in _flatten_stats(PO.make_stats()))
od = torch.from_numpy(offs).to(dev)
out = torch.empty(B, T, F, device=dev)
vp = lambda t: C.c_void_p(t.data_ptr())


def step():
    _lib.check(lib.ishara_preprocess(vp(raw), vp(od), B, int(plens.max()), vp(pm), vp(ps), T, 1, vp(out), None))


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
by = float(offs[-1]) * F * 4 + B * T * F * 4
print(json.dumps({"preprocess_ms": ms, "GBps": by / ms / 1e6, "seq_per_s": B / ms * 1e3}))
