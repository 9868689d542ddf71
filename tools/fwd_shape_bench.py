"""Device-resident forward timing at another frame count (the notebook's INPUT_SHAPE is [176, 276]; BASELINE's is 384).
usage: python tools/fwd_shape_bench.py [T] [B]"""
import sys

import torch

sys.path.insert(0, ".")
import ishara_b200 as ib

T = int(sys.argv[1]) if len(sys.argv) > 1 else 176
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
dev = torch.device("cuda", 0)
m = ib.get_model(input_shape=(T, 276), seed=3)
xs = [torch.randn(B, T, 276, device=dev) for _ in range(4)]
lg = torch.empty(B, T, 60, device=dev)
for i in range(5):
    m.forward_into(xs[i % 4], lg)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
K = 20
for i in range(K):
    m.forward_into(xs[i % 4], lg)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print(f"T={T} B={B}: {ms:.3f} ms per forward, {B / ms * 1e3:.0f} seq/s, {B * T / ms * 1e-3:.1f} M frames/s")
prof = {}
for e in m.profile_forward(xs[0], lg):
    prof[e["label"]] = prof.get(e["label"], 0.0) + e["ms"]
print("  " + ", ".join(f"{k} {v * 1e3:.0f}us" for k, v in sorted(prof.items(), key=lambda kv: -kv[1])[:8]))
