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
