"""Data-parallel training step timing (one process per GPU: torchrun or RANK / WORLD_SIZE / MASTER_* in the env; also runs
with a single process). Same workload as the `train` leg of bench.py. usage: torchrun ... tools/train_dp_bench.py [--steps 20]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ishara_b200 as ib  # noqa: E402
from ishara_b200.parallel import DataParallelTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--batch", type=int, default=64)
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    T, F = 384, 276
    m = ib.get_model(device=local, seed=77)
    m.train_config(0.2, seed=1000 + rank)
    m.compile()
    tr = DataParallelTrainer(m)
    xs = [torch.randn(a.batch, T, F, device=dev, generator=torch.Generator(dev).manual_seed(9000 + 100 * rank + i)) for i in range(4)]
    g = torch.Generator().manual_seed(5)
    lab = torch.full((a.batch, 64), 59, dtype=torch.int32)
    for b in range(a.batch):
        n = int(torch.randint(8, 65, (1,), generator=g))
        lab[b, :n] = torch.randint(0, 59, (n,), generator=g, dtype=torch.int32)
    lab = lab.to(dev)
    for i in range(3):
        tr.train_step(xs[i % 4], lab)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        loss = tr.train_step(xs[i % 4], lab, return_loss=(i == a.steps - 1))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({"world": world, "train_ms_per_step": ms, "seq_per_s": world * a.batch / ms * 1e3, "loss": loss,
                          "env": {k: v for k, v in os.environ.items() if k.startswith(("NCCL_", "ISHARA_"))}}))
    tr.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
