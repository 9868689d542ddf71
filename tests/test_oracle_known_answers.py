"""CPU tests: the oracle against every known answer the reference records and against
independent formulations of the same arithmetic. No GPU, no libishara compute calls."""
import numpy as np
import pytest
import torch

from oracle import ishara_oracle as O


def _count(cfg, prefix):
    return sum(int(np.prod(s)) for n, s in O.param_specs(cfg) if n.startswith(prefix))


def test_block_param_counts_match_recorded_model_summaries():
    # nb:conv-squeezeformer-conformer-test c7:out
    cfg = O.Config(expansion_factor=2, num_conv_per_block=0)
    assert _count(cfg, "stem_conv") == 70656
    assert _count(cfg, "stem_bn") == 1024
    assert _count(cfg, "squeezeformer_0.") == 1077280
    assert _count(cfg, "conformer_0.") == 992000
    # nb:conv-squeezeformer-conformer-test-hyper-zoya c7:out (expansion_factor 4 in the Squeezeformer blocks)
    assert _count(O.Config(expansion_factor=4, num_conv_per_block=0), "squeezeformer_0.") == 1872928


def test_model_totals_match_recorded_summaries():
    # both recorded models end in Dense(dim) + Dense(60): 65,792 + 15,420 (ours ends in Dense(2*dim), c7:61)
    cfg = O.Config(num_conv_per_block=0)
    body = sum(_count(cfg, p) for p in ("stem_", "squeezeformer_", "conformer_"))
    assert body + 65792 + 15420 == 4291452
    nt = sum(int(np.prod(s)) for n, s in O.param_specs(cfg) if "moving_" in n)
    assert nt == 1536
    # 4 x Squeezeformer(ef 4) + 4 x Conformer(expand 2)
    c4 = O.Config(num_conv_per_block=0, num_conv_squeeze_blocks=4, num_conv_conform_blocks=4, expansion_factor=4)
    c2 = O.Config(num_conv_per_block=0, num_conv_squeeze_blocks=4, num_conv_conform_blocks=4, expansion_factor=2)
    total = _count(c4, "stem_") + _count(c4, "squeezeformer_") + _count(c2, "conformer_") + 65792 + 15420
    assert total == 11612604


def test_baseline_config_totals():
    assert O.count_params(O.Config()) == (7591096, 13824)
    assert O.count_params(O.Config(dim=384, num_conv_squeeze_blocks=4, num_conv_conform_blocks=4)) == (33383028, 40704)
    k11 = sum(int(np.prod(s)) for n, s in O.param_specs(O.Config()) if n.startswith("convsqueeze_0_1_"))
    assert k11 == 270597


def test_char_map_pinned_by_fallback_constant():
    # c13:22-23: the constant prediction decodes to "2 a-e -aroe"
    assert "".join(O.num_to_char_fn(O.FALLBACK_IDS)) == "2 a-e -aroe"
    assert len(O.CHAR_TO_NUM) == 60 and O.CHAR_TO_NUM["^"] == 59 and O.CHAR_TO_NUM[" "] == 0
    assert O.CHAR_TO_NUM["a"] == 32 and O.CHAR_TO_NUM["z"] == 57 and O.CHAR_TO_NUM["~"] == 58
    assert O.num_to_char_fn([60, -1]) == ["", ""]


def test_tflite_postprocess_shape_and_fallback():
    # nb:...-hyper-zoya c13:out: TensorShape([13, 59]) for a 13-token prediction
    assert O.tflite_postprocess(np.arange(13)).shape == (13, 59)
    fb = O.tflite_postprocess(np.array([5, 6]))
    assert fb.shape == (11, 59) and list(fb.argmax(1)) == list(O.FALLBACK_IDS)


def test_decode_quirk_final_run_is_dropped():
    def onehot(ids, V=60):
        p = np.zeros((len(ids), V), np.float32)
        p[np.arange(len(ids)), ids] = 1
        return p

    assert list(O.decode_phrase(onehot([3, 3, 4, 4]))) == [3]                 # [a,a,b,b] -> "a"
    assert list(O.decode_phrase(onehot([3, 59, 3, 59]))) == [3, 3]            # blank separates repeats
    assert list(O.decode_phrase(onehot([59, 59, 59]))) == []
    assert list(O.decode_phrase(onehot([7]))) == []                           # T = 1: nothing can be kept
    assert list(O.decode_phrase(onehot([1, 2, 3, 59]))) == [1, 2, 3]
    ties = np.zeros((3, 60), np.float32)                                       # all-equal rows: argmax = 0
    assert list(O.decode_phrase(ties)) == []
    assert O.decode_batch_predictions(np.stack([onehot([32, 32, 33, 59])]))[0] == "ab"


def test_positional_encoding_layout():
    pe = O.positional_encoding(384, 256)
    assert pe.shape == (384, 256) and pe.dtype == np.float32
    assert np.all(pe[0, :128] == 0) and np.all(pe[0, 128:] == 1)               # [sin | cos] halves
    assert abs(pe[1, 0] - np.sin(1.0)) < 1e-6 and abs(pe[1, 128] - np.cos(1.0)) < 1e-6
    assert abs(pe[5, 127] - np.sin(5.0 / 10000 ** (127 / 128))) < 1e-6


def test_forward_two_formulations_agree():
    cfg = O.Config(dim=64, num_conv_squeeze_blocks=1, num_conv_conform_blocks=1, num_heads=4, frames=40,
                   features=20, num_classes=12, kernel_sizes=(5, 3), num_conv_per_block=2, transformer_kernel_size=7)
    p = O.init_params(cfg, seed=3)
    x = O.make_inputs(cfg, 2, ragged=True)
    a = O.forward(p, x, cfg, "float64")
    b = O.forward_np(p, x, cfg)
    assert a.shape == (2, 40, 12)
    assert np.abs(a - b).max() < 1e-10
    c = O.forward(p, x, cfg, "float32")
    assert np.abs(a - c).max() < 1e-4


def test_forward_output_shape_and_mask_modes():
    cfg = O.Config(frames=48)
    p = O.init_params(cfg)
    x = O.make_inputs(cfg, 1)
    lg = O.forward(p, x, cfg)
    assert lg.shape == (1, 48, 60) and lg.dtype == np.float32 and np.isfinite(lg).all()
    # dense input => both mask modes agree exactly
    assert np.array_equal(lg, O.forward(p, x, cfg, mask_mode="propagated"))
    xr = O.make_inputs(cfg, 1, ragged=True)
    xr[0, 30:] = 0
    assert not np.allclose(O.forward(p, xr, cfg), O.forward(p, xr, cfg, mask_mode="propagated"))


def test_forward_is_causal_inside_conv1dblocks_only():
    # Conformer's depthwise conv is 'same' (non-causal) and attention is global, so the whole model is not causal;
    # but the stem is strictly per-frame: perturbing frame t changes stem rows only at t.
    cfg = O.Config(frames=32, num_conv_squeeze_blocks=0, num_conv_conform_blocks=0)
    p = O.init_params(cfg)
    x = O.make_inputs(cfg, 1)
    x2 = x.copy()
    x2[0, 10] += 1
    d = np.abs(O.forward(p, x, cfg) - O.forward(p, x2, cfg)).max(-1)[0]
    assert d[10] > 0 and d[:10].max() == 0 and d[11:].max() == 0


@pytest.mark.parametrize("T,V,L", [(384, 60, 64), (50, 12, 7)])
def test_ctc_matches_torch(T, V, L):
    rng = np.random.default_rng(0)
    B = 4
    logits = (rng.standard_normal((B, T, V)) * 2).astype(np.float32)
    labels = np.full((B, L), V - 1, np.int32)
    lens = [L, 0, L // 2, 3]
    for b, n in enumerate(lens):
        labels[b, :n] = rng.integers(0, V - 1, n)
    labels[0, 1] = labels[0, 0]                                                # repeated label
    nll, grad = O.ctc_loss(labels, logits, blank=V - 1, with_grad=True)
    lg = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    ref = torch.nn.functional.ctc_loss(torch.log_softmax(lg, -1).transpose(0, 1), torch.tensor(labels).long(),
                                       torch.full((B,), T), torch.tensor(lens), blank=V - 1, reduction="none")
    ref.sum().backward()
    assert np.allclose(nll, ref.detach().numpy(), rtol=1e-10, atol=1e-9)
    assert np.abs(grad - lg.grad.numpy()).max() < 1e-9
    assert abs(O.ctc_loss_mean(labels, logits, V - 1) - float(ref.mean())) < 1e-9


def test_ctc_brute_force_tiny():
    rng = np.random.default_rng(1)
    T, V = 5, 4
    logits = rng.standard_normal((3, T, V))
    labels = np.array([[0, 1, 3], [2, 2, 3], [3, 3, 3]], np.int32)             # blank = 3; row 1 has a repeat
    nll = O.ctc_loss(labels, logits, blank=3)
    for b in range(3):
        assert abs(nll[b] - O.ctc_brute_force(labels[b], logits[b], 3)) < 1e-9


def test_ctc_infeasible_is_inf():
    logits = np.zeros((1, 3, 5))
    labels = np.array([[0, 0, 1, 4]], np.int32)                                # needs >= 4 frames (repeat)
    nll = O.ctc_loss(labels, logits, blank=4)
    assert np.isinf(nll[0]) and nll[0] > 0


def test_untrained_loss_magnitude_matches_reference_logs():
    # recorded epoch-1 losses are O(10^2) at T=176, L<=64; an untrained model must land there
    cfg = O.Config(frames=176)
    p = O.init_params(cfg)
    lg = O.forward(p, O.make_inputs(cfg, 2), cfg)
    loss = O.ctc_loss_mean(O.make_labels(cfg, 2), lg)
    assert 30 < loss < 2000


def test_keras_layer_table_covers_every_parameter_once():
    """ishara_b200/keras_names.py: the explicit Keras-layer -> canonical-name table (used to load model.h5 files, c9:10, and
    by tools/dump_tf_reference.py) names every tensor of get_model exactly once, for the BASELINE shapes and the notebook's
    own 4 + 4 call (c7:75-81); the per-layer counts reproduce the recorded model.summary() totals."""
    from ishara_b200.keras_names import keras_layer_table

    for ns, nc in ((2, 2), (4, 4), (1, 0)):
        cfg = O.Config(num_conv_squeeze_blocks=ns, num_conv_conform_blocks=nc)
        specs = dict(O.param_specs(cfg))
        names = [n for _, tw, ntw in keras_layer_table(ns, nc, cfg.num_conv_per_block) for n in tw + ntw]
        assert len(names) == len(set(names)) == len(specs) and set(names) == set(specs)
    table = {l: (tw, ntw) for l, tw, ntw in keras_layer_table()}
    specs = dict(O.param_specs(O.Config()))
    count = lambda ns_: sum(int(np.prod(specs[n])) for n in ns_)
    assert count(table["squeezeformer_0"][0]) == 1_077_280          # nb:conv-squeezeformer-conformer-test c7:out
    assert count(table["conformer_0"][0]) + count(table["conformer_0"][1]) == 992_000
    # Keras order inside a SqueezeformerBlock starts with norm1 and ends with ffn2 (c5:161-180); a ConformerBlock lists its
    # two LayerNorms LAST (c5:314-319)
    assert table["squeezeformer_0"][0][0] == "squeezeformer_0.norm1.gamma" and table["squeezeformer_0"][0][-1] == "squeezeformer_0.ffn2.2.bias"
    assert table["conformer_0"][0][-4:] == ["conformer_0.layer_norm1.gamma", "conformer_0.layer_norm1.beta",
                                            "conformer_0.layer_norm2.gamma", "conformer_0.layer_norm2.beta"]


def test_bf16_emulated_forward_is_the_same_function_up_to_storage_rounding():
    """forward_bf16_emulated = forward + bf16 roundings at the CUDA path's storage points: it must stay within bf16-level
    distance of the fp64 forward (both mask modes) and be exactly reproducible."""
    import numpy as np
    from oracle import ishara_oracle as O
    cfg = O.Config(dim=128, num_heads=4, frames=64, features=20, num_conv_squeeze_blocks=1, num_conv_conform_blocks=1)
    p = O.init_params(cfg, seed=3)
    for mode, ragged in (("dropped", False), ("propagated", True)):
        x = O.make_inputs(cfg, 2, seed=5, ragged=ragged)
        ref = O.forward(p, x, cfg, "float64", mask_mode=mode)
        emu = O.forward_bf16_emulated(p, x, cfg, mask_mode=mode)
        assert np.array_equal(emu, O.forward_bf16_emulated(p, x, cfg, mask_mode=mode))
        rel = np.abs(emu - ref).max() / np.abs(ref).max()
        assert 1e-5 < rel < 3e-2, rel          # not identical (roundings are really applied), not a different function
    # the rounding helper: nearest-even to 8 significant bits
    t = O.torch.tensor([1.0, 1.00390625, 1.005859375, 1.01171875, -3.1415926], dtype=O.torch.float64)
    assert O._r(t).tolist() == [1.0, 1.0, 1.0078125, 1.015625, -3.140625]
