"""GPU landmark preprocessing (ishara_preprocess) against the CPU oracle: same float32 formulas, so the results agree to
rounding (measured: bit-identical except where the compiler's division differs by one ulp) — tolerance 2e-6 relative.
Edge cases: empty sequence, shorter than / equal to / longer than frame_len, all-NaN hands, filter off."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import ishara_b200  # noqa: E402
from oracle import ishara_preprocess_oracle as P  # noqa: E402


@pytest.mark.parametrize("frame_len", [384, 176])
@pytest.mark.parametrize("filter_frames", [True, False])
def test_matches_oracle(frame_len, filter_frames):
    st = P.make_stats()
    lens = [0, 1, 2, 37, frame_len - 1, frame_len, frame_len + 1, 2 * frame_len + 13, 1500]
    seqs = [P.make_frames(n, seed=100 + i) for i, n in enumerate(lens)]
    seqs[3][:, :42] = np.nan                      # right+left hand x gone everywhere in one sequence
    pp = ishara_b200.LandmarkPreprocessor(st, frame_len=frame_len, filter_frames=filter_frames)
    got = pp(seqs)
    assert got.shape == (len(lens), frame_len, 276) and got.dtype == np.float32
    for i, s in enumerate(seqs):
        ref = P.preprocess(s, st, frame_len, filter_frames=filter_frames)
        assert not np.isnan(got[i]).any()
        assert np.allclose(got[i], ref, rtol=2e-6, atol=1e-6), (i, lens[i], float(np.abs(got[i] - ref).max()))
        assert np.array_equal(got[i] == 0, ref == 0)          # the zero (padding / NaN) pattern is exact


def test_feeds_the_model():
    """preprocess -> model(x) end to end: the drop-in for TFLiteModel.__call__ up to decode (c13:9-18)."""
    from oracle import ishara_oracle as O

    cfg = O.Config(frames=176)
    params = O.init_params(cfg)
    st = P.make_stats()
    seqs = [P.make_frames(n, seed=n) for n in (90, 400)]
    x = ishara_b200.LandmarkPreprocessor(st, frame_len=176)(seqs)
    m = ishara_b200.get_model(input_shape=(176, 276)).load_weights(params)
    ids = [ishara_b200.decode_phrase(l) for l in m(x)]
    ref = O.forward(params, np.stack([P.preprocess(s, st, 176) for s in seqs]), cfg)
    assert [i.tolist() for i in ids] == [O.decode_phrase(l).tolist() for l in m(x)]
    assert np.abs(m(x) - ref).max() <= 3e-2 * np.abs(ref).max()
    m.close()


def test_tflite_wrapper_and_fallback():
    """TFLiteModel.__call__ (c13:1-25) end to end, incl. the short-prediction fallback string."""
    from oracle import ishara_oracle as O

    cfg = O.Config(frames=176)
    params = O.init_params(cfg)
    st = P.make_stats()
    frames = P.make_frames(300, seed=3)
    m = ishara_b200.get_model(input_shape=(176, 276)).load_weights(params)
    pre = ishara_b200.LandmarkPreprocessor(st, frame_len=176)
    tfl = ishara_b200.TFLiteModel(m, pre)
    out = tfl(frames)["outputs"]
    ids = ishara_b200.decode_phrase(m(pre([frames]))[0])
    assert np.array_equal(out, ishara_b200.tflite_postprocess(ids))
    assert out.shape[1] == 59 and np.all(out.sum(axis=1) == 1)
    # a model that only ever predicts the pad token decodes to nothing -> constant fallback prediction (c13:21-23)
    p2 = dict(params)
    b = p2["classifier.bias"].copy()
    b[59] += 1e4
    p2["classifier.bias"] = b
    m.load_weights(p2)
    assert tfl.predict_str(frames) == "2 a-e -aroe"
    assert tfl(np.zeros((0, 276), np.float32))["outputs"].shape == (11, 59)     # empty input guard (c13:11)
    m.close()
