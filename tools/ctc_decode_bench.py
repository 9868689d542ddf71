"""CUDA-event timing of the CTC-loss and greedy-decode launches alone at the bench shape (B=256, T=384, V=60, L=64)."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from ishara_b200 import _lib

lib = _lib.load()
B, T, V, L = 256, 384, 60, 64
dev = torch.device("cuda", 0)
g = torch.Generator(dev).manual_seed(1)
logits = [torch.randn(B, T, V, device=dev, generator=g) for _ in range(4)]
labels = torch.full((B, L), V - 1, dtype=torch.int32, device=dev)
for b in range(B):
    n = 8 + (b * 7) % 56
    labels[b, :n] = torch.randint(0, V - 1, (n,), device=dev, generator=g, dtype=torch.int32)
nll = torch.empty(B, device=dev)
ids = torch.empty(B, T, dtype=torch.int32, device=dev)
lens = torch.empty(B, dtype=torch.int32, device=dev)
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
vp = lambda t: C.c_void_p(t.data_ptr())


def timeit(fn, n=40):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


ctc = lambda i: _lib.check(lib.ishara_ctc_loss(vp(logits[i % 4]), vp(labels), B, T, V, L, V - 1, vp(nll), None, sp))
dec = lambda i: _lib.check(lib.ishara_greedy_decode(vp(logits[i % 4]), B, T, V, V - 1, vp(ids), vp(lens), sp))
print("ctc_loss us/launch", round(timeit(ctc), 1), " greedy_decode us/launch", round(timeit(dec), 1))

# training shape: loss + gradient at B = 64
Bt = 64
grad = torch.empty(Bt, T, V, device=dev)
ctcg = lambda i: _lib.check(lib.ishara_ctc_loss(vp(logits[i % 4][:Bt]), vp(labels[:Bt]), Bt, T, V, L, V - 1, vp(nll[:Bt]), vp(grad), sp))
print("ctc_loss + gradient (B=64) us/call", round(timeit(ctcg), 1))
