"""Data-parallel training step on 2 GPUs: the library's own bucketed, overlapped NCCL exchange
(ishara_model_comm_init), identical AdamW update on every rank. Skipped on boxes with a single GPU (the CPU tests cover
the host logic and the bucket plan)."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import ishara_b200
        from ishara_b200.parallel import DataParallelTrainer
        from oracle import ishara_oracle as O

        cfg = O.Config(dim=128, num_heads=4, frames=128, features=20, num_conv_squeeze_blocks=1, num_conv_conform_blocks=1)
        p = O.init_params(cfg)
        m = ishara_b200.get_model(dim=128, num_heads=4, num_conv_squeeze_blocks=1, num_conv_conform_blocks=1, dropout_rate=0.0,
                                  input_shape=(128, 20), device=rank)
        m.load_weights(p)
        m.train_config(0.0)
        m.compile(lr=1e-3, weight_decay=0.0)
        x = O.make_inputs(cfg, 8)
        y = O.make_labels(cfg, 8, max_len=24, min_len=6)
        lo, hi = rank * 4, rank * 4 + 4
        xt, yt = torch.from_numpy(x[lo:hi]).cuda(), torch.from_numpy(y[lo:hi]).cuda()
        # local gradient of this rank's shard, before any exchange
        m.forward_backward(xt, yt)
        g_local = torch.from_dlpack(m.grad_tensor()).clone()
        parts = [torch.empty_like(g_local) for _ in range(world)]
        dist.all_gather(parts, g_local)
        want = sum(parts) / world
        local_loss = torch.tensor([m.last_loss()], device="cuda")
        dist.all_reduce(local_loss)
        tr = DataParallelTrainer(m)               # comm_init inside the library (NCCL id broadcast over the torch group)
        m.train_config(0.0)                       # the trainer seeds dropout per rank; this test runs without dropout
        m.forward_backward_async(xt, yt)          # same data: the buckets are summed over the ranks while backward runs
        g_after = torch.from_dlpack(m.grad_tensor()).clone()
        err = float((g_after / world - want).abs().max() / want.abs().max())
        loss_err = abs(m.last_loss() - float(local_loss.item()) / world)
        assert loss_err < 1e-4 * abs(m.last_loss()), loss_err   # the returned loss is the mean over the ranks
        m.apply_gradients(1.0 / world, m.last_stream)
        losses = [m.last_loss()] + [tr.train_step(xt, yt) for _ in range(5)]
        w = m.get_weights()
        flat = np.concatenate([w[k].ravel() for k in sorted(w) if not k.endswith(("moving_mean", "moving_variance"))])
        digest = torch.from_numpy(flat).cuda()
        allw = [torch.empty_like(digest) for _ in range(world)]
        dist.all_gather(allw, digest)
        same = bool(torch.equal(allw[0], allw[1]))
        if rank == 0:
            q.put((err, losses, same))
    finally:
        dist.destroy_process_group()


def test_two_gpu_data_parallel_step():
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    err, losses, same = q.get()
    assert err < 1e-3, err                      # all-reduced buffer = sum of the per-rank gradients (atomics reorder bits)
    assert same                                 # replicas stay bit-identical after 6 updates
    assert losses[-1] < losses[0], losses
