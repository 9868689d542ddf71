"""CPU tests of the N>1 path (world_size 2, gloo): shard boundaries, rank-ordered gathering and the max-over-ranks
timing reduction that bench.py uses. The per-rank 'model' here is the CPU oracle's decode/CTC on given logits — the GPU
kernels are not involved; what is tested is that sharding + gathering reproduces the unsharded result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ishara_b200.parallel import ShardedInference, gather_in_rank_order, max_over_ranks, shard_range
from oracle import ishara_oracle as O


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 13, 256, 512):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


class _OracleModel:
    """Stands in for IsharaModel.infer on a rank without a GPU: logits are given, decode + CTC by the oracle."""

    def infer(self, x, labels=None):
        ids = [O.decode_phrase(l) for l in x]
        return {"ids": ids, "text": ["".join(O.num_to_char_fn(i)) for i in ids],
                "nll": O.ctc_loss(labels, x).astype(np.float32) if labels is not None else None}


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)                      # same global batch on every rank
        logits = rng.standard_normal((7, 40, 60)).astype(np.float32)
        logits[:, :, 59] += 1.0
        labels = O.make_labels(O.Config(frames=40), 7, max_len=12, min_len=2)
        res = ShardedInference(_OracleModel()).infer(logits, labels)
        slowest = max_over_ranks(1.0 + rank)
        order = gather_in_rank_order([rank] * (rank + 1))
        if rank == 0:
            q.put((res["text"], [i.tolist() for i in res["ids"]], res["nll"].tolist(), slowest, order))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_matches_unsharded():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    text, ids, nll, slowest, order = q.get()
    rng = np.random.default_rng(0)
    logits = rng.standard_normal((7, 40, 60)).astype(np.float32)
    logits[:, :, 59] += 1.0
    labels = O.make_labels(O.Config(frames=40), 7, max_len=12, min_len=2)
    assert text == O.decode_batch_predictions(logits)
    assert ids == [O.decode_phrase(l).tolist() for l in logits]
    assert np.allclose(nll, O.ctc_loss(labels, logits), rtol=1e-6)
    assert slowest == 2.0                                   # max over ranks, not the local value
    assert order == [0, 1, 1]


# ---- data-parallel training step: one all-reduce over the flat gradient buffer -----------------------
class _FakeTrainModel:
    """Host stand-in for IsharaModel's training surface: 'gradient' = mean of the rank's shard, SGD update."""

    def __init__(self):
        self.w = torch.zeros(4, dtype=torch.float32)
        self.g = torch.zeros(4, dtype=torch.float32)
        self.scales = []

    def forward_backward(self, x, labels):
        self.g.copy_(torch.from_numpy(x.mean(axis=0)))
        return float(x.sum())

    def grad_tensor(self):
        return self.g

    def apply_gradients(self, grad_scale=1.0):
        self.scales.append(grad_scale)
        self.w -= self.g * grad_scale


def _train_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ishara_b200.parallel import DataParallelTrainer

        x = np.arange(24, dtype=np.float32).reshape(6, 4)
        lo, hi = shard_range(6, rank, world)
        model = _FakeTrainModel()
        loss = DataParallelTrainer(model).train_step(x[lo:hi], None)
        q.put((rank, model.w.tolist(), loss, model.scales))
    finally:
        dist.destroy_process_group()


def test_data_parallel_trainer_world_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get() for _ in range(2))
    x = np.arange(24, dtype=np.float32).reshape(6, 4)
    want_w = (-x.mean(axis=0)).tolist()                     # equal shards: mean of shard means = global mean
    for rank, w, loss, scales in got:
        assert np.allclose(w, want_w)                        # every rank applied the identical averaged gradient
        assert scales == [0.5]
        assert np.isclose(loss, x.sum() / 2)                 # mean over ranks of the per-rank losses
