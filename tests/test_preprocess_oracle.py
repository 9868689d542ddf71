"""CPU checks of the preprocessing oracle: column bookkeeping against the reference's own list
definitions, the bilinear rule against torch's half-pixel interpolation, padding / filter / empty-input edge cases."""
import numpy as np
import torch

from ishara_b200.preprocess import GROUPS, _flatten_stats, sel_cols
from oracle import ishara_preprocess_oracle as P


def test_sel_cols_and_group_indices_follow_the_reference_lists():
    cols = P.sel_cols()
    assert cols == sel_cols() and len(cols) == 276
    assert cols[0] == "x_right_hand_0" and cols[21] == "x_left_hand_0" and cols[42] == "x_pose_13" and cols[52] == "x_face_61"
    gi = P.group_indices()
    assert [gi[g].shape[0] for g, _ in GROUPS] == [40, 21, 21, 5, 5]
    assert gi["rpose"][:, 0].tolist() == [47, 48, 49, 50, 51] and gi["lpose"][:, 0].tolist() == [42, 43, 44, 45, 46]
    assert gi["lip"][0].tolist() == [52, 144, 236]
    # every selected column is used exactly once by the five groups
    used = np.concatenate([gi[g].ravel() for g, _ in GROUPS])
    assert sorted(used.tolist()) == list(range(276))


def test_bilinear_matches_torch_half_pixel():
    x = np.random.default_rng(0).standard_normal((700, 5, 3)).astype(np.float32)
    for out_len in (384, 176, 699):
        a = P.resize_time_bilinear(x, out_len)
        b = torch.nn.functional.interpolate(torch.from_numpy(x).permute(2, 0, 1)[None], size=(out_len, 5), mode="bilinear",
                                            align_corners=False)[0].permute(1, 2, 0).numpy()
        # same rule; torch rounds the source coordinate in a different order, worth ~1 ulp of 700 = 6e-5 in the weight
        assert np.abs(a - b).max() < 5e-4
    assert np.array_equal(P.resize_time_bilinear(x, 700), x)      # same length = identity


def test_padding_filter_and_empty_input():
    st = P.make_stats()
    mean, std = _flatten_stats(st)
    # short sequence without NaNs, filter off: first N rows are the normalised frames in output column order, rest zero
    x = np.random.default_rng(1).uniform(0, 1, (10, 276)).astype(np.float32)
    y = P.preprocess(x, st, 32, filter_frames=False)
    gi = P.group_indices()
    src = np.concatenate([gi[g].reshape(-1) for g, _ in GROUPS])
    assert np.allclose(y[:10], (x[:, src] - mean) / std, rtol=1e-6)
    assert not y[10:].any()
    # frames whose hands are missing survive only at even indices
    x2 = x.copy()
    x2[:, np.concatenate([gi["rhand"].ravel(), gi["lhand"].ravel()])] = np.nan
    assert P.frame_filter(x2).tolist() == [True, False] * 5
    # empty input = one all-zero frame (c13:11)
    y0 = P.preprocess(np.zeros((0, 276), np.float32), st, 16)
    assert np.allclose(y0[0], (0 - mean) / std, rtol=1e-6) and not y0[1:].any()
