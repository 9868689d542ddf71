"""Multi-GPU plumbing for the hot path: one process per GPU, sequences sharded across ranks.

Inference, loss and decode have no cross-sequence dependency (BatchNorm uses moving statistics), so the path shards by
sequence with NO data-path collective; `torch.distributed` is used only to agree on shard boundaries,
to gather the small results (decoded strings, per-sequence losses) and to reduce timings (max over ranks). The backend
is NCCL on GPUs and gloo in the CPU tests; nothing here touches the kernels.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of `n` sequences owned by `rank`; the first n % world ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        return None
    return dist


def max_over_ranks(value: float, device=None) -> float:
    """Timing reduction the bench contract asks for: the slowest rank defines the step time."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return float(value)
    import torch

    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    if t.is_cuda:
        t = t.float()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_in_rank_order(local: Sequence) -> List:
    """Concatenate per-rank result lists in rank order (= global sequence order under shard_range)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return list(local)
    parts: List[Optional[list]] = [None] * dist.get_world_size()
    dist.all_gather_object(parts, list(local))
    out: List = []
    for p in parts:
        out.extend(p)
    return out


class ShardedInference:
    """Drop-in for `model.infer` on a GLOBAL batch: every rank passes the same global arrays (or only its shard with
    `already_sharded=True`), runs its contiguous slice on its own GPU and gets the global, ordered results back."""

    def __init__(self, model, rank: Optional[int] = None, world: Optional[int] = None):
        dist = _dist()
        self.model = model
        self.rank = rank if rank is not None else (dist.get_rank() if dist else 0)
        self.world = world if world is not None else (dist.get_world_size() if dist else 1)

    def infer(self, x: np.ndarray, labels: Optional[np.ndarray] = None, already_sharded: bool = False) -> dict:
        if not already_sharded:
            lo, hi = shard_range(x.shape[0], self.rank, self.world)
            x = x[lo:hi]
            labels = labels[lo:hi] if labels is not None else None
        if x.shape[0] > 0:
            r = self.model.infer(x, labels=labels)
            text, ids = r["text"], [i.tolist() for i in r["ids"]]
            nll = r["nll"].tolist() if r["nll"] is not None else []
        else:
            text, ids, nll = [], [], []
        return {"text": gather_in_rank_order(text), "ids": [np.asarray(i, np.int64) for i in gather_in_rank_order(ids)],
                "nll": np.asarray(gather_in_rank_order(nll), np.float32) if labels is not None else None}


def allreduce_sum_(t) -> None:
    """In-place SUM all-reduce of a torch tensor over the default process group (NCCL on GPUs, gloo in CPU tests)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return
    dist.all_reduce(t, op=dist.ReduceOp.SUM)


def _env_rank_world() -> Tuple[int, int]:
    import os

    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def exchange_unique_id(make_id, rank: int, world: int, addr: Optional[str] = None, port: Optional[int] = None,
                       timeout: float = 120.0) -> bytes:
    """Rank 0 calls ``make_id()`` (128 bytes: the NCCL unique id) and hands it to every other rank; returns the id on all
    ranks. No PyTorch involved: an initialised ``torch.distributed`` group is used when there is one, otherwise a plain TCP
    rendezvous on (MASTER_ADDR, MASTER_PORT + 17) - the variables torchrun exports."""
    import os
    import socket
    import time

    if world == 1:
        return make_id()
    try:
        dist = _dist()
    except ImportError:
        dist = None
    if dist is not None and dist.get_world_size() == world:
        box = [make_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return bytes(box[0])
    addr = addr or os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = port if port is not None else int(os.environ.get("MASTER_PORT", "29500")) + 17
    if rank == 0:
        uid = make_id()
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind((addr, port))
        srv.listen(world)
        srv.settimeout(timeout)
        try:
            for _ in range(world - 1):
                conn, _peer = srv.accept()
                with conn:
                    conn.sendall(uid)
        finally:
            srv.close()
        return uid
    deadline = time.time() + timeout
    while True:
        try:
            with socket.create_connection((addr, port), timeout=5.0) as c:
                buf = b""
                while len(buf) < 128:
                    chunk = c.recv(128 - len(buf))
                    if not chunk:
                        break
                    buf += chunk
            if len(buf) == 128:
                return buf
        except OSError:
            pass
        if time.time() > deadline:
            raise TimeoutError(f"rank {rank}: no unique id from rank 0 at {addr}:{port}")
        time.sleep(0.05)


def bucket_plan(hi: Sequence[int], n_train: int, min_elems: int = 1 << 20) -> List[Tuple[int, int]]:
    """The library's exchange plan (``ishara_comm_bucket_plan``, pure host logic): for every module k, in forward order,
    the [lo, up) range of the flat gradient buffer that is all-reduced right after module k's backward."""
    import ctypes as C

    from . import _lib

    n = len(hi)
    arr = (C.c_int64 * n)(*[int(v) for v in hi])
    lo, up = (C.c_int64 * n)(), (C.c_int64 * n)()
    _lib.check(_lib.load().ishara_comm_bucket_plan(arr, n, int(n_train), int(min_elems), lo, up))
    return [(int(lo[k]), int(up[k])) for k in range(n)]


class DataParallelTrainer:
    """Data-parallel training step. The exchange lives INSIDE the library: ``model.comm_init`` binds an NCCL
    communicator to the handle, and ``ishara_model_train_forward_backward`` then all-reduces the gradients in per-module
    buckets on a second stream while the backward pass is still running (no host synchronisation, the loss stays on the
    device until the update has been enqueued). Every rank applies the identical AdamW update with grad_scale = 1/world;
    BatchNorm statistics stay per rank, as in the reference (no SyncBN). PyTorch is not needed: rank / world come from
    the arguments or from RANK / WORLD_SIZE, the 128-byte NCCL id travels over ``exchange_unique_id``."""

    def __init__(self, model, rank: Optional[int] = None, world: Optional[int] = None, seed: int = 0):
        if rank is None or world is None:
            try:
                dist = _dist()
            except ImportError:
                dist = None
            if dist is not None:
                rank, world = dist.get_rank(), dist.get_world_size()
            else:
                rank, world = _env_rank_world()
        self.model = model
        self.rank, self.world = int(rank), int(world)
        if self.world > 1:
            uid = exchange_unique_id(model.comm_unique_id, self.rank, self.world)
            model.comm_init(uid, self.rank, self.world)
        # per-rank dropout streams: same seed would give every rank the same masks
        model.train_config(seed=int(seed) + 0x9E3779B1 * self.rank)

    def train_step(self, x, labels, return_loss: bool = True) -> Optional[float]:
        """x/labels = THIS rank's shard. Returns the mean loss over all ranks (reduced on the device by the library), or
        None with ``return_loss=False`` - then nothing synchronises with the host and steps can be enqueued back to back."""
        self.model.forward_backward_async(x, labels)
        self.model.apply_gradients(1.0 / self.world, self.model.last_stream)
        return self.model.last_loss() if return_loss else None

    def close(self):
        if self.world > 1:
            self.model.comm_destroy()
