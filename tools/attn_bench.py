"""Micro-benchmark of the attention core (ishara_op_attention) at the BASELINE shape: B x 384 tokens, 8 heads of 32.
usage: python tools/attn_bench.py [B] [reps]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ishara_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
T, H, dh = 384, 8, 32
lib = _lib.load()
dev = torch.device("cuda", 0)
qkv = (torch.randn(B * T, 3 * H * dh, device=dev) * 0.7).bfloat16()
out = torch.empty(B * T, H * dh, device=dev, dtype=torch.bfloat16)
st = torch.cuda.current_stream(dev)
sp = C.c_void_p(st.cuda_stream)
vp = lambda t: C.c_void_p(t.data_ptr())
for _ in range(3):
    _lib.check(lib.ishara_op_attention(vp(qkv), vp(out), None, B, T, H, dh, 1.0 / 16.0, sp))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(st)
for _ in range(reps):
    _lib.check(lib.ishara_op_attention(vp(qkv), vp(out), None, B, T, H, dh, 1.0 / 16.0, sp))
e1.record(st)
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
# reference check against torch (fp32)
q, k, v = (qkv.float().view(B, T, H, 3, dh)[:, :, :, i].permute(0, 2, 1, 3) for i in range(3))
ref = torch.softmax(q[:4] @ k[:4].transpose(-1, -2) / 16.0, -1) @ v[:4]
got = out.float().view(B, T, H, dh)[:4].permute(0, 2, 1, 3)
err = float((got - ref).abs().max() / ref.abs().max())
print(f"attention B={B}: {us:.1f} us per launch, {4.0 * B * T * T * H * dh / us / 1e6:.1f} TFLOP/s, max rel err {err:.3g}, env stagger={os.environ.get('ISHARA_ATTN_STAGGER')} tc2={os.environ.get('ISHARA_ATTN_TC2')}")
