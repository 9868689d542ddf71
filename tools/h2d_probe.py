"""Host<->device copy bandwidth through the C ABI (pinned via torch, pinned via ishara_host_malloc_pinned, pageable)."""
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ishara_b200 import _lib

lib = _lib.load()
n = 256 * 384 * 276
dev = C.c_void_p()
_lib.check(lib.ishara_device_malloc(0, n * 4, C.byref(dev)))


def bench(name, ptr):
    for _ in range(2):
        _lib.check(lib.ishara_memcpy_async(dev, C.c_void_p(ptr), n * 4, 1, None))
        _lib.check(lib.ishara_stream_synchronize(0, None))
    t0 = time.perf_counter()
    for _ in range(5):
        _lib.check(lib.ishara_memcpy_async(dev, C.c_void_p(ptr), n * 4, 1, None))
    _lib.check(lib.ishara_stream_synchronize(0, None))
    dt = (time.perf_counter() - t0) / 5
    print(f"{name}: {n * 4 / dt / 1e9:.1f} GB/s ({dt * 1e3:.2f} ms for {n * 4 / 1e6:.0f} MB)")


a = torch.randn(n).pin_memory().numpy()
bench("torch pinned", a.ctypes.data)
p = C.c_void_p()
_lib.check(lib.ishara_host_malloc_pinned(n * 4, C.byref(p)))
bench("ishara pinned", p.value)
b = np.random.rand(n).astype(np.float32)
bench("pageable", b.ctypes.data)
t = torch.randn(n).pin_memory()
torch.cuda.synchronize()
d = torch.empty(n, device="cuda")
for _ in range(2):
    d.copy_(t, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    d.copy_(t, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 5
print(f"torch copy_: {n * 4 / dt / 1e9:.1f} GB/s")
