"""Host logic of IsharaModel.infer_pipelined against a stand-in for the C library (no GPU): submission order, at most
two batches in flight, result order, output-slot reuse, and draining when the consumer stops early."""
import ctypes as C

import numpy as np

from ishara_b200 import model as M


class _FakeLib:
    def __init__(self):
        self.inflight, self.submitted, self.collected, self.max_inflight, self.keep = [], 0, 0, 0, []

    def ishara_model_infer_submit(self, h, x, B, lab, L, lg, ids, lens, nll):
        assert len(self.inflight) < 2, "a third batch was submitted before a collect"
        self.inflight.append((B, ids, lens, nll))
        self.submitted += 1
        self.max_inflight = max(self.max_inflight, len(self.inflight))
        return 0

    def ishara_model_infer_collect(self, h):
        B, ids, lens, nll = self.inflight.pop(0)
        self.collected += 1
        np.ctypeslib.as_array(C.cast(ids, C.POINTER(C.c_int32)), shape=(B, 8))[:] = self.collected   # batch serial number
        np.ctypeslib.as_array(C.cast(lens, C.POINTER(C.c_int32)), shape=(B,))[:] = 2
        if nll:
            np.ctypeslib.as_array(C.cast(nll, C.POINTER(C.c_float)), shape=(B,))[:] = 10.0 * self.collected
        return 0

    def ishara_ids_to_text(self, ids, lens, B, T, chars, n, buf, offs):
        np.ctypeslib.as_array(C.cast(offs, C.POINTER(C.c_int64)), shape=(B + 1,))[:] = np.arange(B + 1) * 2
        C.memmove(buf, b"ab" * B, 2 * B)
        return 0

    def ishara_host_malloc_pinned(self, n, out):
        b = (C.c_char * n)()
        self.keep.append(b)
        C.cast(out, C.POINTER(C.c_void_p))[0] = C.addressof(b)
        return 0

    def ishara_host_free_pinned(self, p):
        return 0

    def ishara_model_destroy(self, h):
        return 0


def _model():
    m = object.__new__(M.IsharaModel)
    m._lib, m._h, m._finalized = _FakeLib(), C.c_void_p(1), True
    m.frames, m.features, m.num_classes = 8, 4, 60
    return m


def test_results_come_back_in_order_with_two_in_flight():
    m = _model()
    sizes = (3, 2, 5, 1, 4)
    batches = [(np.zeros((b, 8, 4), np.float32), np.zeros((b, 6), np.int32)) for b in sizes]
    out = list(m.infer_pipelined(batches))
    assert [len(r["text"]) for r in out] == list(sizes)
    assert [int(r["ids"][0][0]) for r in out] == [1, 2, 3, 4, 5]          # oldest batch first
    assert [float(r["nll"][0]) for r in out] == [10.0, 20.0, 30.0, 40.0, 50.0]
    assert all(r["text"] == ["ab"] * b for r, b in zip(out, sizes))
    assert m._lib.max_inflight == 2 and m._lib.submitted == m._lib.collected == 5
    # results were copied out of the reusable output slots
    assert int(out[0]["ids"][0][0]) == 1 and int(out[2]["ids"][0][0]) == 3


def test_early_close_drains_the_pipeline_and_bad_input_raises():
    m = _model()
    batches = [np.zeros((2, 8, 4), np.float32)] * 4
    g = m.infer_pipelined(batches)
    next(g)
    g.close()
    assert m._lib.inflight == [] and m._lib.submitted == m._lib.collected
    try:
        list(m.infer_pipelined([np.zeros((2, 8, 4), np.float32), np.zeros((2, 7, 4), np.float32)]))
        raise AssertionError("shape error expected")
    except ValueError:
        pass
    assert m._lib.inflight == []                                          # the batch in flight was collected
    assert [r["nll"] for r in m.infer_pipelined(batches[:1])] == [None]   # no labels -> no losses


def test_lr_schedule_and_fit_callbacks_host_logic():
    """lrfn / LR_SCHEDULE (c11:1-12,57) and fit()'s per-epoch learning-rate + weight-decay callbacks (c11:59-70, c12:1-9)."""
    import math

    import ishara_b200 as ib

    s = ib.lr_schedule(n_epochs=50, n_warmup_epochs=5, lr_max=4e-3)              # the reference's constants (c10:1-5)
    assert len(s) == 50 and s[:6] == [4e-3 * 2 ** -5, 4e-3 * 2 ** -4, 4e-3 * 2 ** -3, 4e-3 * 2 ** -2, 4e-3 * 2 ** -1, 4e-3]
    assert abs(s[27] - 0.5 * (1 + math.cos(math.pi * 22 / 45)) * 4e-3) < 1e-15 and s[-1] < 1e-5
    assert all(a >= b for a, b in zip(s[5:], s[6:]))                            # monotone decay after the warm-up
    assert abs(ib.lrfn(1, 3, 1.0, warmup_method="log") - 0.01) < 1e-15

    m = _model()
    seen = []
    m.train_step = lambda x, y: (seen.append((m._opt.lr, m._opt.weight_decay)), 2.0 * len(seen))[1]
    m.infer = lambda x, labels=None: {"nll": np.full(len(x), 3.0, np.float32)}
    data = [(np.zeros((2, 8, 4), np.float32), np.zeros((2, 5), np.int32))] * 3
    h = m.fit(data, epochs=2, validation_data=data[:1], lr_schedule=[1e-3, 5e-4], wd_ratio=0.05)
    assert h["loss"] == [4.0, 10.0] and h["val_loss"] == [3.0, 3.0]
    lrs = [round(a, 9) for a, _ in seen]
    assert lrs == [1e-3] * 3 + [5e-4] * 3
    assert all(abs(wd - lr * 0.05) < 1e-9 for lr, wd in seen)                   # WeightDecayCallback: wd = lr * WD_RATIO
