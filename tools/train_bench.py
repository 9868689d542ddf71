"""Times the training step (B=64, T=384, fwd + CTC + bwd + clip + AdamW) on one GPU with CUDA
events; optional short mode for ncu launch lists. python tools/train_bench.py [--batch 64] [--steps 10] [--warmup 3]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ishara_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dropout", type=float, default=0.2)
    a = ap.parse_args()
    torch.cuda.set_device(0)
    m = ishara_b200.get_model(dropout_rate=a.dropout)
    m.train_config(a.dropout, seed=1)
    m.compile()
    rng = np.random.default_rng(0)
    xs = [torch.from_numpy(rng.standard_normal((a.batch, 384, 276)).astype(np.float32)).cuda() for _ in range(4)]
    y = np.full((a.batch, 64), 59, np.int32)
    for b in range(a.batch):
        n = int(rng.integers(8, 65))
        y[b, :n] = rng.integers(0, 59, size=n)
    yt = torch.from_numpy(y).cuda()
    for i in range(a.warmup):
        m.train_step(xs[i % 4], yt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ishara_b200._lib.load().ishara_launch_count()
    e0.record()
    losses = []
    for i in range(a.steps):
        m.forward_backward_async(xs[i % 4], yt) if hasattr(m, "forward_backward_async") else losses.append(m.train_step(xs[i % 4], yt))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    n1 = ishara_b200._lib.load().ishara_launch_count()
    flops = 3 * 6.367e9 * a.batch
    print(json.dumps({"train_ms_per_step": ms, "train_seq_per_s": a.batch / ms * 1e3, "batch": a.batch,
                      "tflops": flops / ms / 1e9, "launches_per_step": (n1 - n0) / a.steps, "dropout": a.dropout,
                      "losses": [round(v, 3) for v in losses[:5]]}))


if __name__ == "__main__":
    main()
