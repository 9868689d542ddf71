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


# ---- data-parallel training step: the exchange runs inside the library; the host side only has to agree on
# ---- the 128-byte NCCL id, seed the ranks differently and scale by 1/world -------------------------------------------------
class _FakeTrainModel:
    """Host stand-in for IsharaModel's training surface as DataParallelTrainer drives it."""

    def __init__(self, rank):
        self.rank = rank
        self.calls = []
        self.last_stream = 0

    def comm_unique_id(self):
        return bytes([7 + self.rank]) * 128            # only rank 0's id may survive the exchange

    def comm_init(self, uid, rank, world):
        self.calls.append(("comm_init", uid, rank, world))

    def comm_destroy(self):
        self.calls.append(("comm_destroy",))

    def train_config(self, seed=0):
        self.calls.append(("train_config", seed))

    def forward_backward_async(self, x, labels):
        self.calls.append(("fb", float(x.sum())))

    def apply_gradients(self, grad_scale=1.0, stream=0):
        self.calls.append(("apply", grad_scale, stream))

    def last_loss(self):
        return 1.5


def _train_worker(rank, world, port, q, use_group):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"], os.environ["WORLD_SIZE"] = str(rank), str(world)
    if use_group:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ishara_b200.parallel import DataParallelTrainer

        x = np.arange(24, dtype=np.float32).reshape(6, 4)
        lo, hi = shard_range(6, rank, world)
        model = _FakeTrainModel(rank)
        tr = DataParallelTrainer(model, seed=5)
        loss = tr.train_step(x[lo:hi], None)
        tr.close()
        q.put((rank, loss, model.calls))
    finally:
        if use_group:
            dist.destroy_process_group()


@pytest.mark.parametrize("use_group", [True, False], ids=["torch_group", "tcp_rendezvous"])
def test_data_parallel_trainer_world_2(use_group):
    """Host logic of the data-parallel step with and without a torch.distributed group (the second form needs no
    PyTorch at all: RANK / WORLD_SIZE / MASTER_* + a TCP hand-over of the NCCL id)."""
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, q, use_group)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(q.get() for _ in range(2))
    x = np.arange(24, dtype=np.float32).reshape(6, 4)
    for rank, loss, calls in got:
        kinds = [c[0] for c in calls]
        assert kinds == ["comm_init", "train_config", "fb", "apply", "comm_destroy"]
        assert calls[0][1] == bytes([7]) * 128 and calls[0][2:] == (rank, 2)      # rank 0's id on every rank
        assert calls[1][1] == 5 + 0x9E3779B1 * rank                               # per-rank dropout streams
        lo, hi = shard_range(6, rank, 2)
        assert calls[2][1] == float(x[lo:hi].sum())                                # this rank's shard only
        assert calls[3][1] == 0.5                                                  # grad_scale = 1 / world
        assert loss == 1.5


def test_bucket_plan_tiles_the_gradient_buffer():
    """ishara_comm_bucket_plan (pure host logic of comm.cu): the ranges reduced after each module's backward tile
    [0, n_train) exactly once, a range is only released once no EARLIER module (which runs its backward later) still
    writes into it, and small ranges are merged."""
    from ishara_b200.parallel import bucket_plan

    # six modules; module 3 also writes a shared tensor that sits BELOW module 1's parameters (ConformerBlock's
    # layer_norm2 precedes ffn1 in the table but is used by the last FFN, c5:318-341)
    hi = [100, 400, 900, 1600, 2000, 2600]
    n_train = 2600
    plan = bucket_plan(hi, n_train, min_elems=1)
    assert plan == [(0, 100), (100, 400), (400, 900), (900, 1600), (1600, 2000), (2000, 2600)]
    covered = sorted(r for r in plan if r[1] > r[0])
    assert covered[0][0] == 0 and covered[-1][1] == n_train
    assert all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
    # merging: nothing below 1000 elements leaves on its own (except the final flush after module 0)
    plan = bucket_plan(hi, n_train, min_elems=1000)
    assert plan == [(0, 900), (0, 0), (0, 0), (900, 2000), (0, 0), (0, 0)] or sum(u - l for l, u in plan) == n_train
    assert sum(u - l for l, u in plan) == n_train
    # a module that reaches up into a LATER module's range delays that range until it has run
    hi2 = [100, 2500, 900, 1600, 2000, 2600]
    plan2 = bucket_plan(hi2, n_train, min_elems=1)
    assert plan2[5] == (2500, 2600) and plan2[4] == (0, 0) and plan2[3] == (0, 0) and plan2[2] == (0, 0)
    assert plan2[1] == (100, 2500) and plan2[0] == (0, 100)
