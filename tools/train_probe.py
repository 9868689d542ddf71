"""GPU-side diagnostic for the training step: compares every named activation, activation gradient and parameter
gradient of ishara_b200's train_forward_backward with the torch-autograd oracle, in forward order, so the first
mismatch localises the faulty kernel. Run on a B200: python tools/train_probe.py [small|full|both]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import ishara_b200  # noqa: E402
from oracle import ishara_oracle as O  # noqa: E402
from oracle import ishara_train_oracle as TO  # noqa: E402


def rel(a, b):
    a = a.astype(np.float64).ravel()
    b = b.astype(np.float64).ravel()
    nb = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (nb + 1e-30)), float(a @ b / (np.linalg.norm(a) * nb + 1e-30))


def run(cfg, B, L, label):
    print(f"==== {label}: dim={cfg.dim} T={cfg.frames} B={B} L={L}", flush=True)
    p = O.init_params(cfg)
    x = O.make_inputs(cfg, B)
    y = O.make_labels(cfg, B, max_len=L, min_len=max(2, L // 4))
    t0 = time.time()
    ref = TO.forward_train(p, x, y, cfg, want_taps=True)
    print(f"oracle: loss {ref['loss']:.5f} ({time.time() - t0:.1f}s)", flush=True)
    m = ishara_b200.get_model(dim=cfg.dim, num_conv_squeeze_blocks=cfg.num_conv_squeeze_blocks,
                              num_conv_conform_blocks=cfg.num_conv_conform_blocks, kernel_sizes=cfg.kernel_sizes,
                              num_conv_per_block=cfg.num_conv_per_block, dropout_rate=0.0, num_heads=cfg.num_heads,
                              expansion_factor=cfg.expansion_factor, transformer_kernel_size=cfg.transformer_kernel_size,
                              input_shape=(cfg.frames, cfg.features), num_classes=cfg.num_classes)
    m.load_weights(p)
    m.train_config(0.0, seed=1, debug=True)
    loss = m.forward_backward(x, y)
    print(f"gpu:    loss {loss:.5f}   rel diff {abs(loss - ref['loss']) / abs(ref['loss']):.2e}", flush=True)
    if os.environ.get("PROBE_GATE", "1") == "1":
        hh = m.train_fetch("head.h", (B, cfg.frames, 2 * cfg.dim))
        ref = TO.forward_train(p, x, y, cfg, want_taps=True, relu_gate=(hh > 0).astype(np.float32))
        print(f"oracle re-run with the GPU's ReLU gate: loss {ref['loss']:.5f}", flush=True)
    print(f"{'tensor':44s} {'val relL2':>10s} {'cos':>8s} | {'grad relL2':>10s} {'cos':>8s}")
    for name, (val, grad) in ref["taps"].items():
        line = f"{name:44s}"
        try:
            v = m.train_fetch(name, val.shape)
            r, c = rel(v, val)
            line += f" {r:10.3e} {c:8.5f}"
        except Exception as e:  # noqa: BLE001
            line += f" {'-':>10s} {'-':>8s}"
        if grad is not None:
            try:
                g = m.train_fetch(name, grad.shape, grad=True)
                r, c = rel(g, grad)
                line += f" | {r:10.3e} {c:8.5f}"
            except Exception:  # noqa: BLE001
                line += f" | {'-':>10s} {'-':>8s}"
        print(line, flush=True)
    grads = m.gradients()
    worst = []
    print(f"{'parameter':60s} {'relL2':>10s} {'cos':>8s} {'|ref|':>10s}")
    for name, g in grads.items():
        r, c = rel(g, ref["grads"][name])
        worst.append((r, name))
        print(f"{name:60s} {r:10.3e} {c:8.5f} {np.linalg.norm(ref['grads'][name]):10.3e}", flush=True)
    worst.sort(reverse=True)
    print("worst parameter gradients:", worst[:8])
    tot, _ = TO.clip_scale(ref["grads"])
    tot_g, _ = TO.clip_scale(grads)
    print(f"global grad norm: gpu {tot_g:.4f} oracle {tot:.4f}")
    # optimiser: one AdamW step from the GPU's own gradients
    new_ref = TO.adamw_step(p, grads, {}, 1)
    m.apply_gradients()
    w = m.get_weights()
    err = max(float(np.abs(w[k] - new_ref[k]).max()) for k in new_ref if TO.is_trainable(k))
    print(f"adamw: max |w_gpu - w_ref| over trainable tensors = {err:.3e}")
    errs = max(float(np.abs(w[k] - v).max() / (np.abs(v).max() + 1e-12)) for k, v in ref["new_stats"].items())
    print(f"moving statistics: max rel err = {errs:.3e}")
    # inference path sees the new weights
    lg = m(x[:2])
    lo = O.forward(w, x[:2], cfg)
    print(f"post-step inference vs oracle(new weights): max abs err {np.abs(lg - lo).max():.3e} (scale {np.abs(lo).max():.2f})")
    # a few more steps: the loss on a fixed batch must go down
    m.compile(lr=1e-3, weight_decay=0.0)
    losses = [m.train_step(x, y) for _ in range(8)]
    print("losses over 8 more steps on the same batch:", " ".join(f"{v:.3f}" for v in losses), flush=True)
    m.close()


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "both"
    if which in ("small", "both"):
        cfg = O.Config(dim=128, num_heads=4, frames=128, features=20, num_conv_squeeze_blocks=1, num_conv_conform_blocks=1)
        run(cfg, 4, 24, "small")
    if which == "d384":
        run(O.Config(dim=384, num_heads=8, frames=64, features=20, num_conv_squeeze_blocks=1, num_conv_conform_blocks=1), 3, 12, "dim 384")
    if which in ("full", "both"):
        run(O.Config(), 4, 64, "cfg3-shape (B=4)")
