"""GPU parity of the operator-level rows P1-P9 against golden vectors produced by RUNNING the reference's
own PyTorch modules (tests/golden/make_torch_ops_golden.py imports them from the reference checkout; the .npz is
committed because the GPU box has no checkout). Tolerance: bf16 activations / fp32 accumulation vs the reference's fp32:
max |err| <= 3e-2 of the output's max magnitude (stated per test where tighter)."""
import os

import numpy as np
import pytest

from ishara_b200.torch_ops import VendoredOps, _sub

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "torch_ops_golden.npz")


@pytest.fixture(scope="module")
def G():
    z = np.load(GOLD)
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="module")
def ops():
    return VendoredOps(0)


def sd(G, prefix):
    return _sub(G, prefix + ".w.")


def close(got, ref, what, rtol=3e-2):
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    scale = np.abs(ref).max()
    err = np.abs(got - ref).max()
    print(f"{what}: max_abs_err={err:.4g} scale={scale:.4g} rel={err / scale:.4g}")
    assert np.isfinite(got).all() and err <= rtol * scale, what


def test_p3_rel_positional_encoding(G):
    B, T, D, H, K = G["meta"]
    tab = VendoredOps.rel_positional_encoding(int(T), int(D))
    assert np.abs(tab - G["p3.pos_emb"][0]).max() < 1e-6


def test_p1_relative_mha_with_and_without_mask(G, ops):
    B, T, D, H, K = (int(v) for v in G["meta"])
    pos = G["p3.pos_emb"][0]
    close(ops.relative_mha(G["x"], sd(G, "p1"), pos, H), G["p1.out"], "P1 rel-pos MHA")
    close(ops.relative_mha(G["x"], sd(G, "p1"), pos, H, mask=G["p1.mask"]), G["p1.out_masked"], "P1 rel-pos MHA masked")


def test_p2_mhsa_module(G, ops):
    H = int(G["meta"][3])
    close(ops.mhsa_module(G["x"], {"attention." + k: v for k, v in sd(G, "p1").items()}, H), G["p2.out"], "P2 MHSA module")


def test_p4_feed_forward(G, ops):
    close(ops.feed_forward(G["x"], sd(G, "p4")), G["p4.out"], "P4 FFN")


def test_p5_conv_module(G, ops):
    close(ops.conv_module(G["x"], sd(G, "p5")), G["p5.out"], "P5 ConvModule")


def test_p6_squeezeformer_block(G, ops):
    H = int(G["meta"][3])
    close(ops.squeezeformer_block(G["x"], sd(G, "p6"), H, half_step_residual=True), G["p6.out"], "P6 SqueezeformerBlock")


def test_p7_time_reduction(G, ops):
    red, lens, proj = ops.time_reduction(G["x"], sd(G, "p7"), np.array([48, 43]), _sub(G, "p7.proj.w."))
    close(red, G["p7.reduced"], "P7 TimeReductionLayer")
    assert list(lens) == list(G["p7.lengths"])
    close(proj, G["p7.out"], "P7 time_reduction_proj")


def test_p8_recover_resolution(G, ops):
    out = ops.recover(G["p7.out"], G["x"], sd(G, "p8"))
    close(out, G["p8.out"], "P8 recover")


def test_p8_conv2d_subsampling(G, ops):
    out, lens = ops.conv2d_subsampling(G["p8b.x"], sd(G, "p8b"), np.array([48, 43]))
    close(out, G["p8b.out"], "P8 DepthwiseConv2dSubsampling")
    assert list(lens) == list(G["p8b.lengths"])


def test_p9_conformer_block(G, ops):
    H = int(G["meta"][3])
    r = ops.conformer_block(G["x"], sd(G, "p9"), H, parts=True)
    close(r["ffn1"], G["p9.ffn1"], "P9 FFN")
    close(r["attn"], G["p9.attn"], "P9 MHSA")
    close(r["conv"], G["p9.conv"], "P9 conv")
    close(r["out"], G["p9.out"], "P9 block")


# ---- second golden set at the BASELINE shape (B=2, T=384, D=256, H=8 -> dh=32, k=15; key mask ending inside a 64-key
# ---- block): the instantiations the toy set above never reaches ------------------------------------------------------
GOLD384 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "torch_ops_golden_t384.npz")


@pytest.fixture(scope="module")
def G384():
    z = np.load(GOLD384)
    return {k: z[k] for k in z.files}


def sd16(G, prefix):
    """weights stored as bf16 bit patterns (the reference ran with exactly these bf16-representable values)"""
    out = {}
    for k, v in G.items():
        if k.startswith(prefix + ".w16."):
            out[k[len(prefix) + 5:]] = (v.astype(np.uint32) << 16).view(np.float32)
    return out


def test_t384_rel_pos_attention_masked_and_module(G384, ops):
    B, T, D, H, K = (int(v) for v in G384["meta"])
    assert (T, D, H) == (384, 256, 8)
    pos = VendoredOps.rel_positional_encoding(T, D)
    w = sd16(G384, "p1")
    close(ops.relative_mha(G384["x"], w, pos, H, mask=G384["p1.mask"]), G384["p1.out_masked"], "P1 rel-pos MHA masked, T=384 dh=32")
    close(ops.mhsa_module(G384["x"], {"attention." + k: v for k, v in w.items()}, H), G384["p2.out"], "P2 MHSA module, T=384")


def test_t384_conv_module_k15(G384, ops):
    close(ops.conv_module(G384["x"], sd16(G384, "p5")), G384["p5.out"], "P5 ConvModule k=15 D=256")


def test_t384_time_reduction(G384, ops):
    red, lens = ops.time_reduction(G384["x"], sd(G384, "p7"), G384["p7.in_lengths"])
    close(red, G384["p7.reduced"], "P7 TimeReductionLayer [384, 256]")
    assert list(lens) == list(G384["p7.lengths"])
