"""Generate golden input/output vectors for the operator-level parity rows P1-P9 by RUNNING THE
REFERENCE ITSELF: the vendored PyTorch files squeezeformer/{attention,convolution,modules,encoder}.py and
conformer/conformer.py are imported from the read-only reference checkout (nothing is copied from them).

`squeezeformer/modules.py:21` imports a module the reference does not ship (`squeezeformer.activation`), so an
in-memory shim providing `Swish` (x * sigmoid(x), the same definition as convolution.py:24-29) is registered and the
package object is created empty so that the broken `__init__.py` -> `model.py` chain is skipped.

Run where /root/reference exists:   python tests/golden/make_torch_ops_golden.py
Writes tests/golden/torch_ops_golden.npz (committed; the GPU box has no reference checkout).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("ISHARA_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "torch_ops_golden.npz")


def load_reference():
    pkg = types.ModuleType("squeezeformer")
    pkg.__path__ = [os.path.join(REF, "squeezeformer")]
    sys.modules["squeezeformer"] = pkg
    act = types.ModuleType("squeezeformer.activation")

    class Swish(nn.Module):
        def forward(self, x):
            return x * x.sigmoid()

    act.Swish = Swish
    sys.modules["squeezeformer.activation"] = act
    mods = {}
    for name in ("modules", "convolution", "attention", "encoder"):
        spec = importlib.util.spec_from_file_location(f"squeezeformer.{name}", os.path.join(REF, "squeezeformer", f"{name}.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"squeezeformer.{name}"] = m
        spec.loader.exec_module(m)
        mods[name] = m
    spec = importlib.util.spec_from_file_location("ref_conformer", os.path.join(REF, "conformer", "conformer.py"))
    cm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cm)
    mods["conformer"] = cm
    return mods


def randomize(module, g):
    """Non-trivial deterministic parameters and BatchNorm running statistics."""
    with torch.no_grad():
        for n, p in module.named_parameters():
            if p.dim() >= 2:
                p.copy_(torch.randn(p.shape, generator=g) * (1.0 / max(1, p.shape[-1] if p.dim() == 2 else p[0].numel())) ** 0.5)
            elif "weight" in n:  # norm scales
                p.copy_(torch.rand(p.shape, generator=g) * 0.4 + 0.8)
            else:
                p.copy_(torch.randn(p.shape, generator=g) * 0.1)
        for n, b in module.named_buffers():
            if n.endswith("running_mean"):
                b.copy_(torch.randn(b.shape, generator=g) * 0.1)
            elif n.endswith("running_var"):
                b.copy_(torch.rand(b.shape, generator=g) + 0.5)


def sd(prefix, module, out):
    for k, v in module.state_dict().items():
        if v.dtype.is_floating_point:
            out[f"{prefix}.w.{k}"] = v.detach().numpy().astype(np.float32)


def main():
    torch.manual_seed(0)
    g = torch.Generator().manual_seed(1234)
    R = load_reference()
    B, T, D, H, K = 2, 48, 64, 4, 7
    out = {"meta": np.array([B, T, D, H, K], np.int64)}
    x = torch.randn(B, T, D, generator=g)
    out["x"] = x.numpy()

    with torch.no_grad():
        # P3 RelPositionalEncoding (modules.py:59-108)
        pe = R["modules"].RelPositionalEncoding(D)
        pos = pe(x)
        out["p3.pos_emb"] = pos.numpy()                                   # [1, 2T-1, D]

        # P1 RelativeMultiHeadAttention (attention.py:25-110)
        att = R["attention"].RelativeMultiHeadAttention(D, H, dropout_p=0.0).eval()
        randomize(att, g)
        with torch.no_grad():
            att.u_bias.copy_(torch.randn(H, D // H, generator=g) * 0.3)
            att.v_bias.copy_(torch.randn(H, D // H, generator=g) * 0.3)
        sd("p1", att, out)
        out["p1.out"] = att(x, x, x, pos.repeat(B, 1, 1)).numpy()
        mask = torch.zeros(B, 1, T, dtype=torch.bool)
        mask[1, 0, T - 9:] = True                                         # True = masked (attention.py:92-94)
        out["p1.mask"] = mask.numpy()
        out["p1.out_masked"] = att(x, x, x, pos.repeat(B, 1, 1), mask=mask).numpy()

        # P2 MultiHeadedSelfAttentionModule (attention.py:113-139) — same weights, PE generated inside
        mod = R["attention"].MultiHeadedSelfAttentionModule(D, H, dropout_p=0.0).eval()
        mod.attention.load_state_dict(att.state_dict())
        out["p2.out"] = mod(x).numpy()

        # P4 FeedForwardModule (modules.py:24-56)
        ffn = R["modules"].FeedForwardModule(D, expansion_factor=4, dropout_p=0.0).eval()
        randomize(ffn, g)
        sd("p4", ffn, out)
        out["p4.out"] = ffn(x).numpy()

        # P5 ConvModule (convolution.py:199-238), inference BatchNorm
        conv = R["convolution"].ConvModule(D, kernel_size=K, expansion_factor=2, dropout_p=0.0).eval()
        randomize(conv, g)
        sd("p5", conv, out)
        out["p5.out"] = conv(x).numpy()

        # P6 SqueezeformerBlock (encoder.py:169-247), half-step residual, post-LN
        blk = R["encoder"].SqueezeformerBlock(encoder_dim=D, num_attention_heads=H, feed_forward_expansion_factor=4,
                                              conv_expansion_factor=2, feed_forward_dropout_p=0.0, attention_dropout_p=0.0,
                                              conv_dropout_p=0.0, conv_kernel_size=K, half_step_residual=True).eval()
        randomize(blk, g)
        with torch.no_grad():
            a = blk.sequential[0].module.attention
            a.u_bias.copy_(torch.randn(H, D // H, generator=g) * 0.3)
            a.v_bias.copy_(torch.randn(H, D // H, generator=g) * 0.3)
        sd("p6", blk, out)
        out["p6.out"] = blk(x).numpy()

        # P7 TimeReductionLayer (convolution.py:241-269) + time_reduction_proj (encoder.py:80,155)
        tr = R["convolution"].TimeReductionLayer().eval()
        randomize(tr, g)
        proj = nn.Linear((D - 1) // 2, D)
        randomize(proj, g)
        sd("p7", tr, out)
        sd("p7.proj", proj, out)
        lens = torch.tensor([T, T - 5])
        red, rl = tr(x, lens.clone())
        out["p7.reduced"] = red.numpy()                                   # [B, (T-3)//2+1, (D-3)//2+1]
        out["p7.lengths"] = rl.numpy()
        out["p7.out"] = proj(red).numpy()

        # P8 recover_resolution (modules.py:137-142) + recover step (encoder.py:157-162)
        rec = nn.Linear(D, D)
        randomize(rec, g)
        sd("p8", rec, out)
        small = proj(red)                                                  # [B, 23, D]
        up = R["modules"].recover_resolution(small)
        out["p8.upsampled"] = up.numpy()
        y = rec(up)
        y = y + x[:, : up.size(1), :]
        out["p8.out"] = y.numpy()

        # P8b DepthwiseConv2dSubsampling (convolution.py:39-73)
        sub = R["convolution"].DepthwiseConv2dSubsampling(1, 8).eval()
        randomize(sub, g)
        sd("p8b", sub, out)
        xin = torch.randn(B, T, 20, generator=g)
        out["p8b.x"] = xin.numpy()
        so, sl = sub(xin, lens.clone())
        out["p8b.out"] = so.numpy()
        out["p8b.lengths"] = sl.numpy()

        # P9 conformer/conformer.py ConformerBlock (:59-73) with its FFN (:6-22), MHSA (:24-35), conv (:37-57)
        cb = R["conformer"].ConformerBlock(D, num_heads=H, expansion_factor=4, kernel_size=K, dropout=0.0).eval()
        randomize(cb, g)
        sd("p9", cb, out)
        out["p9.ffn1"] = cb.ffn1(x).numpy()
        out["p9.attn"] = cb.attention(x).numpy()
        out["p9.conv"] = cb.conv(x).numpy()
        out["p9.out"] = cb(x).numpy()

    np.savez_compressed(OUT, **out)
    print("wrote", OUT, f"{os.path.getsize(OUT) / 1e3:.0f} KB,", len(out), "arrays")


def bf16_round(t):
    """Round to bf16-representable fp32 values (stored as uint16 halves in the .npz to keep the fixture small): the CUDA
    path rounds weights to bf16 anyway, so the reference run and the kernels see IDENTICAL weights."""
    return t.to(torch.bfloat16).to(torch.float32)


def main_t384():
    """Second set at the BASELINE shape: B=2, T=384, D=256, H=8 (dh=32), k=15 and a key mask that ends in the MIDDLE of a
    64-key block (284 = 4*64 + 28 valid keys). Only the rows whose kernels change instantiation with the shape: P1/P2
    (rel-pos attention at dh=32, T=384), P5 (ConvModule k=15, C=512), P7 (time reduction of a [384, 256] plane)."""
    g = torch.Generator().manual_seed(4321)
    R = load_reference()
    B, T, D, H, K = 2, 384, 256, 8, 15
    out = {"meta": np.array([B, T, D, H, K], np.int64)}
    x = torch.randn(B, T, D, generator=g)
    out["x"] = x.numpy()

    def sd16(prefix, module):
        for k, v in module.state_dict().items():
            if v.dtype.is_floating_point:
                out[f"{prefix}.w16.{k}"] = v.detach().to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)

    def round_params(module):
        with torch.no_grad():
            for p_ in module.parameters():
                p_.copy_(bf16_round(p_))
            for b_ in module.buffers():
                if b_.dtype.is_floating_point:
                    b_.copy_(bf16_round(b_))

    with torch.no_grad():
        pe = R["modules"].RelPositionalEncoding(D)
        pos = pe(x)
        att = R["attention"].RelativeMultiHeadAttention(D, H, dropout_p=0.0).eval()
        randomize(att, g)
        att.u_bias.copy_(torch.randn(H, D // H, generator=g) * 0.3)
        att.v_bias.copy_(torch.randn(H, D // H, generator=g) * 0.3)
        round_params(att)
        sd16("p1", att)
        mask = torch.zeros(B, 1, T, dtype=torch.bool)
        mask[1, 0, 284:] = True                                           # True = masked; ends mid-block (284 = 4*64 + 28)
        mask[0, 0, 383:] = True
        out["p1.mask"] = mask.numpy()
        out["p1.out_masked"] = att(x, x, x, pos.repeat(B, 1, 1), mask=mask).numpy()
        mod = R["attention"].MultiHeadedSelfAttentionModule(D, H, dropout_p=0.0).eval()
        mod.attention.load_state_dict(att.state_dict())
        out["p2.out"] = mod(x).numpy()

        conv = R["convolution"].ConvModule(D, kernel_size=K, expansion_factor=2, dropout_p=0.0).eval()
        randomize(conv, g)
        round_params(conv)
        sd16("p5", conv)
        out["p5.out"] = conv(x).numpy()

        tr = R["convolution"].TimeReductionLayer().eval()
        randomize(tr, g)
        sd("p7", tr, out)                                                  # 10 numbers: keep fp32
        lens = torch.tensor([T, T - 37])
        red, rl = tr(x, lens.clone())
        out["p7.reduced"] = red.numpy()
        out["p7.lengths"] = rl.numpy()
        out["p7.in_lengths"] = lens.numpy()

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "torch_ops_golden_t384.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, f"{os.path.getsize(path) / 1e3:.0f} KB,", len(out), "arrays")


if __name__ == "__main__":
    main()
    main_t384()
