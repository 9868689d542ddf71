"""CPU checks of the training-step oracle and of the host-side training logic: autograd gradients
against central finite differences in float64, AdamW + clip against torch.optim.AdamW / clip_grad_norm_, BatchNorm
moving-statistics rule, and the host replication of the kernels' dropout masks."""
import numpy as np
import torch

import ishara_b200
from oracle import ishara_oracle as O
from oracle import ishara_train_oracle as TO

TINY = O.Config(dim=64, num_heads=2, frames=48, features=12, num_conv_squeeze_blocks=1, num_conv_conform_blocks=1)


def _setup(B=2, L=8):
    p = O.init_params(TINY, round_bf16=False)
    return p, O.make_inputs(TINY, B), O.make_labels(TINY, B, max_len=L, min_len=3)


def test_autograd_matches_finite_differences_float64():
    p, x, y = _setup()
    p = {k: v.astype(np.float64) for k, v in p.items()}
    r = TO.forward_train(p, x, y, TINY, dtype="float64")
    rng = np.random.default_rng(0)
    for name in ("stem_conv.kernel", "convsqueeze_0_2_dwconv.depthwise_kernel", "convsqueeze_0_1_eca.kernel",
                 "squeezeformer_0.mha.qkv.kernel", "squeezeformer_0.conv.se.fc1.kernel", "conformer_0.conv.batch_norm.gamma",
                 "conformer_0.conv.depthwise_conv.kernel", "conformer_0.conv.layer_norm.beta", "classifier.bias"):
        g = r["grads"][name].astype(np.float64)
        for _ in range(2):
            idx = tuple(rng.integers(0, s) for s in p[name].shape)
            eps = 1e-5
            pp, pm = dict(p), dict(p)
            a = p[name].copy(); a[idx] += eps; pp[name] = a
            b = p[name].copy(); b[idx] -= eps; pm[name] = b
            fd = (TO.forward_train(pp, x, y, TINY, dtype="float64")["loss"] - TO.forward_train(pm, x, y, TINY, dtype="float64")["loss"]) / (2 * eps)
            assert abs(fd - g[idx]) <= 1e-5 * max(1.0, abs(fd)), (name, idx, fd, g[idx])


def test_loss_is_the_mean_ctc_of_the_training_mode_logits():
    p, x, y = _setup()
    r = TO.forward_train(p, x, y, TINY)
    assert np.isclose(r["loss"], float(np.mean(O.ctc_loss(y, r["logits"]))), rtol=1e-5)
    assert np.allclose(r["nll"], O.ctc_loss(y, r["logits"]), rtol=1e-5)
    # batch statistics, not moving statistics: the logits differ from the inference forward
    assert np.abs(r["logits"] - O.forward(p, x, TINY)).max() > 1e-3


def test_adamw_and_clip_match_torch():
    p, x, y = _setup()
    r = TO.forward_train(p, x, y, TINY)
    names = [k for k in p if TO.is_trainable(k)]
    tp = {k: torch.nn.Parameter(torch.from_numpy(p[k].copy())) for k in names}
    opt = torch.optim.AdamW(tp.values(), lr=4.5e-3, weight_decay=0.08)
    state, cur = {}, dict(p)
    for step in (1, 2, 3):
        for k in names:
            tp[k].grad = torch.from_numpy(r["grads"][k].copy())
        total = torch.nn.utils.clip_grad_norm_(tp.values(), 1.0)
        opt.step()
        cur = TO.adamw_step(cur, r["grads"], state, step)
        assert np.isclose(float(total), TO.clip_scale(r["grads"])[0], rtol=1e-5)
        for k in names:
            assert np.abs(cur[k] - tp[k].detach().numpy()).max() <= 1e-6, (step, k)   # a couple of fp32 ulps at |w| ~ 1


def test_moving_statistics_rule():
    p, x, y = _setup()
    r = TO.forward_train(p, x, y, TINY, want_taps=True)
    z = r["taps"]["stem.z"][0]
    want = 0.95 * p["stem_bn.moving_mean"] + 0.05 * z.mean(axis=(0, 1))  # BatchNormalization(momentum=0.95, name='stem_bn') c7:17
    assert np.allclose(r["new_stats"]["stem_bn.moving_mean"], want, atol=1e-6)
    d = r["taps"]["convsqueeze_0_1.d"][0]
    want_v = 0.95 * p["convsqueeze_0_1_bn.moving_variance"] + 0.05 * d.var(axis=(0, 1))
    assert np.allclose(r["new_stats"]["convsqueeze_0_1_bn.moving_variance"], want_v, atol=1e-6)


def test_host_dropout_masks_structure():
    """IsharaModel.dropout_masks is pure host code (the kernels' counter-based hash restated in numpy)."""
    m = ishara_b200.get_model(dim=128, num_heads=4, input_shape=(64, 20), num_conv_squeeze_blocks=1, num_conv_conform_blocks=1)
    a = m.dropout_masks(3, seed=11, rate=0.25)
    b = m.dropout_masks(3, seed=11, rate=0.25)
    c = m.dropout_masks(3, seed=12, rate=0.25)
    assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)
    assert any(not np.array_equal(a[k], c[k]) for k in a)
    assert m.dropout_masks(3, seed=11, rate=0.0) == {}
    # sites of the reference: Conv1DBlock per-sample, FFN inner (both block types), Squeezeformer residual branches,
    # attention probabilities, head 0.4; ConformerBlock has no residual-branch dropout
    assert a["convsqueeze_0_1.drop"].shape == (3, 1, 1)
    assert a["squeezeformer_0.ffn1.drop"].shape == (3, 64, 256) and a["squeezeformer_0.drop2"].shape == (3, 64, 128)
    assert a["squeezeformer_0.mha.attn_drop"].shape == (3, 4, 64, 64)
    assert "conformer_0.drop1" not in a and "conformer_0.ffn2.drop" in a
    for k, v in a.items():
        rate = 0.4 if k == "head.drop" else (0.1 if k == "conformer_0.mha.attn_drop" else 0.25)
        nz = v[v != 0]
        assert nz.size == 0 or np.allclose(nz, 1 / (1 - rate), rtol=1e-6), k
        if v.size > 1000:
            assert abs(float((v == 0).mean()) - rate) < 0.02, k
    m.close()


def test_oracle_against_tf_training_golden_if_present():
    """tools/dump_tf_reference.py (run where TensorFlow is available) pins the training-mode oracle: loss, gradients
    and the BatchNorm moving-statistics update of the real Keras model with dropout switched off."""
    import os

    import pytest

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tf_reference_train.npz")
    if not os.path.exists(path):
        pytest.skip("no TensorFlow golden vectors (parity unpinned; see tools/dump_tf_reference.py)")
    z = np.load(path, allow_pickle=True)
    p = {k[2:]: z[k] for k in z.files if k.startswith("w:")}
    cfg = O.Config(frames=int(z["x"].shape[1]))
    r = TO.forward_train(p, z["x"], z["labels"], cfg)
    assert abs(r["loss"] - float(z["loss"])) <= 1e-4 * abs(float(z["loss"]))
    assert np.abs(r["logits"] - z["logits_train"]).max() <= 1e-3 * np.abs(z["logits_train"]).max()
    for k in z.files:
        if k.startswith("g:"):
            ref = z[k]
            assert np.linalg.norm(r["grads"][k[2:]] - ref) <= 2e-3 * max(np.linalg.norm(ref), 1e-6), k
        if k.startswith("s:"):
            assert np.allclose(r["new_stats"][k[2:]], z[k], rtol=1e-4, atol=1e-6), k


def test_radam_matches_torch_and_lookahead_rule():
    """RectifiedAdam restatement (tfa is absent: third-party, unpinned) vs torch.optim.RAdam on the steps where the two
    rectification thresholds agree (tfa: sma_t >= threshold, torch: rho_t > 5; with threshold 5 they agree at every step),
    then the Lookahead rule on top of it."""
    import torch

    rng = np.random.default_rng(3)
    p0 = {"w": rng.standard_normal((7, 5)).astype(np.float32), "b": rng.standard_normal(5).astype(np.float32)}
    grads = [{k: rng.standard_normal(v.shape).astype(np.float32) for k, v in p0.items()} for _ in range(12)]
    tp = {k: torch.tensor(v.astype(np.float64), requires_grad=True) for k, v in p0.items()}
    opt = torch.optim.RAdam(list(tp.values()), lr=2e-3, betas=(0.9, 0.999), eps=1e-7, weight_decay=0.0)
    p, state = dict(p0), {}
    for t, g in enumerate(grads, 1):
        for k in tp:
            tp[k].grad = torch.tensor(g[k].astype(np.float64))
        opt.step()
        p = TO.radam_lookahead_step(p, g, state, t, lr=2e-3, eps=1e-7, sma_threshold=5.0, sync_period=0)
        for k in p:
            assert np.allclose(p[k], tp[k].detach().numpy(), rtol=2e-6, atol=1e-7), (t, k)
    # the reference's setting (threshold 4, c7:68): step 5 is already rectified (sma_5 = 4.98), steps 1-4 are momentum SGD
    p, state = dict(p0), {}
    prev = None
    for t, g in enumerate(grads, 1):
        before = {k: v.copy() for k, v in p.items()}
        p = TO.radam_lookahead_step(p, g, state, t, lr=2e-3, sma_threshold=4.0, sync_period=5)
        if t <= 4:   # un-rectified: theta -= lr * mhat
            mhat = state["w"]["m"] / (1 - 0.9 ** t)
            assert np.allclose(p["w"], before["w"] - 2e-3 * mhat, atol=1e-7)
        if t % 5 == 0:  # Lookahead sync: fast == slow afterwards
            assert np.allclose(p["w"], state["w"]["slow"], atol=0)
            if prev is not None:
                pass
    # slow weights after the first sync = initial + 0.5 * (fast_5 - initial)
    p, state = dict(p0), {}
    for t, g in enumerate(grads[:5], 1):
        fast_before_sync = TO.radam_lookahead_step(p, g, {k: {kk: vv.copy() for kk, vv in v.items()} for k, v in state.items()}, t,
                                                   lr=2e-3, sync_period=0)
        p = TO.radam_lookahead_step(p, g, state, t, lr=2e-3, sync_period=5)
    assert np.allclose(p["w"], p0["w"] + 0.5 * (fast_before_sync["w"] - p0["w"]), atol=1e-7)
