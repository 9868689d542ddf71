"""CPU oracle for ONE TRAINING STEP of the Ishara encoder — TEST INFRASTRUCTURE ONLY.

Same rules as ``ishara_oracle.py``: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may
import this file; ``ishara_b200`` never does. PARITY STATUS: **parity unpinned** (TensorFlow/Keras unavailable; the
reference ships no recorded gradients). The gradients come from torch autograd over a float32/float64 restatement of
the Keras training-mode forward, so the only hand-written arithmetic is the forward itself.

What it restates (``nb:conv-hybrid-model cN:L``):
  * ``model(x, training=True)`` c7:12-65 with c5:1-343: BatchNormalization uses the biased batch statistics over
    (B, T) and updates ``moving = m*moving + (1-m)*batch`` (m = 0.95 for ``Conv1DBlock`` c5:73 and ``stem_bn`` c7:17, Keras default
    0.99 for ``ConvolutionModule.batch_norm`` c5:281); dropout sites: FFN inner + residual branches
    (c5:162-166,183,190,204), attention probabilities (c5:113), per-sample ``noise_shape=(None,1,1)`` on the
    Conv1DBlock branch (c5:83), head 0.4 (c7:62). Dropout is driven by explicit keep-masks (or off) so the CUDA path
    can be compared deterministically.
  * loss = ``CTCLoss`` c6:1-13 = mean over the batch of ``tf.nn.ctc_loss`` (blank = 59, logit_length = T).
  * optimiser per BASELINE.json: AdamW (lr 4.5e-3, weight_decay 0.08, betas (0.9, 0.999), eps 1e-8, decoupled decay
    on every trainable tensor) after global-norm clipping at 1.0 — ``integration.py:675-679,750``.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .ishara_oracle import (BN_EPS, LN_EPS, LN_EPS_CONVMOD, Config, _dwconv, _swish, param_specs, positional_encoding)

BN_MOMENTUM_CONV1D = 0.95   # c5:73
BN_MOMENTUM_STEM = 0.95     # BatchNormalization(momentum=0.95, name='stem_bn') c7:17
BN_MOMENTUM_DEFAULT = 0.99  # Keras default: ConvolutionModule.batch_norm (c5:281)


def is_trainable(name: str) -> bool:
    return not (name.endswith(".moving_mean") or name.endswith(".moving_variance"))


class _Ctx:
    """Parameters as autograd leaves + taps of named intermediates (with retained gradients)."""

    def __init__(self, params: Dict[str, np.ndarray], dtype, want_taps: bool):
        self.dt = dtype
        self.p = {}
        for k, v in params.items():
            t = torch.from_numpy(np.ascontiguousarray(v)).to(dtype)
            if is_trainable(k):
                t.requires_grad_(True)
            self.p[k] = t
        self.taps: Dict[str, torch.Tensor] = {}
        self.want_taps = want_taps
        self.new_stats: Dict[str, torch.Tensor] = {}

    def tap(self, name, t):
        if self.want_taps:
            if t.requires_grad:
                t.retain_grad()
            self.taps[name] = t
        return t

    def dense(self, x, base, bias=True):
        w = self.p[base + ".kernel"]
        if w.dim() == 3:
            w = w[0]
        y = x @ w
        return y + self.p[base + ".bias"] if bias else y

    def ln(self, x, base, eps):
        return F.layer_norm(x, (x.shape[-1],), self.p[base + ".gamma"], self.p[base + ".beta"], eps)

    def bn_train(self, x, base, momentum):
        mu = x.mean(dim=(0, 1))
        var = x.var(dim=(0, 1), unbiased=False)
        with torch.no_grad():
            self.new_stats[base + ".moving_mean"] = momentum * self.p[base + ".moving_mean"] + (1 - momentum) * mu
            self.new_stats[base + ".moving_variance"] = momentum * self.p[base + ".moving_variance"] + (1 - momentum) * var
        return (x - mu) / torch.sqrt(var + BN_EPS) * self.p[base + ".gamma"] + self.p[base + ".beta"]


def _drop(x, mask):
    return x if mask is None else x * mask


def forward_train(params: Dict[str, np.ndarray], x: np.ndarray, labels: np.ndarray, cfg: Config,
                  dtype: str = "float32", dropout_masks: Optional[Dict[str, np.ndarray]] = None,
                  want_taps: bool = False, relu_gate: Optional[np.ndarray] = None):
    """One training-mode forward + backward. Returns dict(loss, nll[B], grads{name: np}, new_stats{name: np},
    logits, taps{name: (value, grad)}). ``dropout_masks[name]`` are ready-to-multiply masks (keep/(1-p)), broadcastable
    to the tensor they scale; missing name = dropout off at that site. ``relu_gate`` (0/1, shape of the head's hidden
    tensor) replaces the head's ReLU by a fixed gate: a bf16 forward flips the sign of the ~1 % of pre-activations that
    sit within rounding distance of zero, and each flip is an O(1) gradient difference, so backward kernels are
    compared "given the same gate" (the un-gated comparison is reported too)."""
    dt = {"float32": torch.float32, "float64": torch.float64}[dtype]
    c = _Ctx(params, dt, want_taps)
    dm = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dt) for k, v in (dropout_masks or {}).items()}
    D, H, tk = cfg.dim, cfg.num_heads, cfg.transformer_kernel_size
    xt = torch.from_numpy(np.ascontiguousarray(x)).to(dt)
    B, T, _ = xt.shape

    def ffn(h, base):
        u = c.tap(base + ".u", c.dense(h, base + ".0"))
        return c.dense(_drop(_swish(u), dm.get(base + ".drop")), base + ".2")

    def mhsa(h, base):
        qkv = c.tap(base + ".qkv", c.dense(h, base + ".qkv", bias=False))
        q4 = qkv.view(B, T, H, 3 * D // H).permute(0, 2, 1, 3)
        q, k, v = torch.split(q4, D // H, dim=-1)
        attn = torch.softmax((q @ k.transpose(-1, -2)) * (D ** -0.5), dim=-1)
        attn = _drop(attn, dm.get(base + ".attn_drop"))
        o = c.tap(base + ".o", (attn @ v).permute(0, 2, 1, 3).reshape(B, T, D))
        return c.dense(o, base + ".proj", bias=False)

    def conv_blocks(h, tag, i):
        for j in range(cfg.num_conv_per_block):
            k = cfg.kernel_sizes[j % len(cfg.kernel_sizes)]
            n = f"conv{tag}_{i}_{j + 1}"
            skip = h
            e = c.tap(n + ".e", c.dense(h, n + "_expand_conv"))
            d = c.tap(n + ".d", _dwconv(_swish(e), c.p[n + "_dwconv.depthwise_kernel"][:, :, 0], k - 1, 0))
            bn = c.bn_train(d, n + "_bn", BN_MOMENTUM_CONV1D)
            m = bn.mean(dim=1)
            s = torch.sigmoid(F.conv1d(m.unsqueeze(1), c.p[n + "_eca.kernel"].view(1, 1, 5), padding=2).squeeze(1))
            g = c.tap(n + ".g", bn * s.unsqueeze(1))
            y = c.dense(g, n + "_project_conv")
            h = c.tap(n, skip + _drop(y, dm.get(n + ".drop")))
        return h

    h = c.dense(xt, "stem_conv", bias=False) + torch.from_numpy(positional_encoding(T, D)).to(dt)
    h = c.tap("stem.z", h)
    h = c.tap("stem", c.bn_train(h, "stem_bn", BN_MOMENTUM_STEM))
    for i in range(cfg.num_conv_squeeze_blocks):
        h = conv_blocks(h, "squeeze", i)
        n = f"squeezeformer_{i}"
        h = c.tap(n + ".x1", h + _drop(ffn(c.ln(h, n + ".norm1", LN_EPS), n + ".ffn1"), dm.get(n + ".drop1")))
        h = c.tap(n + ".x2", h + _drop(mhsa(c.ln(h, n + ".norm2", LN_EPS), n + ".mha"), dm.get(n + ".drop2")))
        u = c.tap(n + ".conv.c1", c.dense(c.ln(h, n + ".conv.norm", LN_EPS), n + ".conv.conv1"))
        d2 = c.tap(n + ".conv.d2", _dwconv(_swish(u), c.p[n + ".conv.conv2.depthwise_kernel"][:, :, 0], tk - 1, 0))
        z = c.tap(n + ".conv.z", c.dense(_swish(d2), n + ".conv.conv3"))
        g = _swish(c.dense(z.mean(dim=1), n + ".conv.se.fc1"))
        g = torch.sigmoid(c.dense(g, n + ".conv.se.fc2"))
        h = c.tap(n + ".x3", z * g.unsqueeze(1) + h)
        h = c.tap(n, h + _drop(ffn(c.ln(h, n + ".norm3", LN_EPS), n + ".ffn2"), dm.get(n + ".drop3")))
    for i in range(cfg.num_conv_conform_blocks):
        h = conv_blocks(h, "conform", i)
        n = f"conformer_{i}"
        h = c.tap(n + ".x1", h + _drop(ffn(c.ln(h, n + ".layer_norm1", LN_EPS), n + ".ffn1"), dm.get(n + ".drop1")))
        h = c.tap(n + ".x2", h + _drop(mhsa(c.ln(h, n + ".layer_norm1", LN_EPS), n + ".mha"), dm.get(n + ".drop2")))
        res = h
        p1 = c.tap(n + ".conv.p1", c.dense(h, n + ".conv.pointwise_conv1"))
        gl = c.tap(n + ".conv.gl", p1[..., :D] * torch.sigmoid(p1[..., D:]))
        w = c.p[n + ".conv.depthwise_conv.kernel"][:, 0, :]
        dw = c.tap(n + ".conv.dw", _dwconv(gl, w, (tk - 1) // 2, tk - 1 - (tk - 1) // 2, c.p[n + ".conv.depthwise_conv.bias"]))
        bn = c.tap(n + ".conv.bn", c.bn_train(dw, n + ".conv.batch_norm", BN_MOMENTUM_DEFAULT))
        r = c.tap(n + ".conv.r", c.dense(bn, n + ".conv.pointwise_conv2") + res)
        h = c.tap(n + ".x3", c.ln(r, n + ".conv.layer_norm", LN_EPS_CONVMOD))
        h = c.tap(n, h + _drop(ffn(c.ln(h, n + ".layer_norm2", LN_EPS), n + ".ffn2"), dm.get(n + ".drop3")))
    pre = c.dense(h, "top_conv")
    if relu_gate is None:
        hh = c.tap("head.h", torch.relu(pre))
    else:
        hh = c.tap("head.h", pre * torch.from_numpy(np.ascontiguousarray(relu_gate)).to(dt))
    logits = c.tap("logits", c.dense(_drop(hh, dm.get("head.drop")), "classifier"))

    # CTCLoss c6:1-13: label_length = count(labels != pad), logit_length = T, blank = pad index, mean over batch
    lab = torch.from_numpy(np.ascontiguousarray(labels)).long()
    lab_len = (lab != cfg.blank).sum(dim=1)
    lp = torch.log_softmax(logits, dim=-1).transpose(0, 1)  # [T,B,V]
    nll = F.ctc_loss(lp, lab, torch.full((B,), T, dtype=torch.long), lab_len, blank=cfg.blank, reduction="none",
                     zero_infinity=False)
    loss = nll.mean()
    loss.backward()
    grads = {k: (v.grad.detach().float().numpy() if v.grad is not None else np.zeros(tuple(v.shape), np.float32))
             for k, v in c.p.items() if is_trainable(k)}
    taps = {}
    for k, v in c.taps.items():
        taps[k] = (v.detach().float().numpy(), None if v.grad is None else v.grad.detach().float().numpy())
    return dict(loss=float(loss.detach()), nll=nll.detach().float().numpy(), grads=grads,
                new_stats={k: v.float().numpy() for k, v in c.new_stats.items()},
                logits=logits.detach().float().numpy(), taps=taps)


def clip_scale(grads: Dict[str, np.ndarray], max_norm: float = 1.0) -> Tuple[float, float]:
    """torch.nn.utils.clip_grad_norm_ (integration.py:750): returns (total_norm, scale)."""
    total = float(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads.values())))
    return total, min(1.0, max_norm / (total + 1e-6))


def adamw_step(params: Dict[str, np.ndarray], grads: Dict[str, np.ndarray], state: Dict[str, Dict[str, np.ndarray]],
               step: int, lr: float = 4.5e-3, weight_decay: float = 0.08, beta1: float = 0.9, beta2: float = 0.999,
               eps: float = 1e-8, max_norm: float = 1.0) -> Dict[str, np.ndarray]:
    """torch.optim.AdamW semantics (decoupled decay, bias-corrected moments) after global-norm clipping.
    ``step`` is 1-based. Returns the new parameter dict; ``state`` (exp_avg / exp_avg_sq) is updated in place."""
    _, scale = clip_scale(grads, max_norm) if max_norm and max_norm > 0 else (0.0, 1.0)
    out = dict(params)
    for k, g in grads.items():
        g = g.astype(np.float64) * scale
        st = state.setdefault(k, {"m": np.zeros_like(g), "v": np.zeros_like(g)})
        st["m"] = beta1 * st["m"] + (1 - beta1) * g
        st["v"] = beta2 * st["v"] + (1 - beta2) * g * g
        p = params[k].astype(np.float64) * (1 - lr * weight_decay)
        denom = np.sqrt(st["v"]) / np.sqrt(1 - beta2 ** step) + eps
        out[k] = (p - (lr / (1 - beta1 ** step)) * st["m"] / denom).astype(np.float32)
    return out


def radam_lookahead_step(params: Dict[str, np.ndarray], grads: Dict[str, np.ndarray], state: Dict[str, Dict[str, np.ndarray]],
                         step: int, lr: float = 1e-3, weight_decay: float = 0.0, beta1: float = 0.9, beta2: float = 0.999,
                         eps: float = 1e-7, sma_threshold: float = 4.0, sync_period: int = 5, slow_step_size: float = 0.5,
                         max_norm: float = 0.0) -> Dict[str, np.ndarray]:
    """The optimiser the reference compiles the model with (c7:68-69):
    ``tfa.optimizers.Lookahead(tfa.optimizers.RectifiedAdam(sma_threshold=4), sync_period=5)``.

    tensorflow_addons is a third-party dependency that is neither vendored nor pinned (``Dockerfile:20-21``) and cannot be
    imported here, so this restates its published algorithm (``tensorflow_addons/optimizers/rectified_adam.py`` with
    ``total_steps=0`` -> no warm-up schedule; ``lookahead.py`` with ``slow_step_size=0.5``): bias-corrected Adam moments,
    the variance-rectification term r_t once the SMA length reaches ``sma_threshold`` (plain momentum SGD before),
    decoupled weight decay folded into the update, and every ``sync_period`` steps the slow weights move half way to the
    fast ones and the fast ones are reset to them. The RectifiedAdam part is cross-checked against ``torch.optim.RAdam``
    (threshold 5) in tests/test_train_oracle.py. ``step`` is 1-based; ``state`` holds m / v / slow per tensor."""
    _, scale = clip_scale(grads, max_norm) if max_norm and max_norm > 0 else (0.0, 1.0)
    sma_inf = 2.0 / (1.0 - beta2) - 1.0
    b2t = beta2 ** step
    sma_t = sma_inf - 2.0 * step * b2t / (1.0 - b2t)
    rect = None
    if sma_t >= sma_threshold:
        rect = np.sqrt((sma_t - 4.0) / (sma_inf - 4.0) * (sma_t - 2.0) / (sma_inf - 2.0) * sma_inf / sma_t)
    out = dict(params)
    for k, g in grads.items():
        g = g.astype(np.float64) * scale
        p = params[k].astype(np.float64)
        st = state.setdefault(k, {"m": np.zeros_like(g), "v": np.zeros_like(g), "slow": p.copy()})
        st["m"] = beta1 * st["m"] + (1 - beta1) * g
        st["v"] = beta2 * st["v"] + (1 - beta2) * g * g
        mhat = st["m"] / (1 - beta1 ** step)
        upd = rect * mhat / (np.sqrt(st["v"] / (1 - b2t)) + eps) if rect is not None else mhat
        upd = upd + weight_decay * p
        p = p - lr * upd
        if sync_period > 0 and step % sync_period == 0:
            st["slow"] = st["slow"] + slow_step_size * (p - st["slow"])
            p = st["slow"].copy()
        out[k] = p.astype(np.float32)
    return out


def train_step(params, x, labels, cfg, state, step, dtype="float32", **adamw):
    r = forward_train(params, x, labels, cfg, dtype=dtype)
    new = adamw_step(params, r["grads"], state, step, **adamw)
    new.update(r["new_stats"])
    return r["loss"], new
