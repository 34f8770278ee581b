"""Property tests of the oracle (CPU only): SURVEY.md section 8(c) items (1)-(7).

hypothesis draws the shapes and the data (few distinct values => many exact ties, planted duplicates);
the C restatement, the numpy restatement and -- when the wheel is importable -- OpenCV's own
batchDistance / BFMatcher must agree bit for bit, and the structural properties the consumers rely on
(/root/reference/src/Sfm.cpp:549-553, 700-711: ascending queryIdx, one entry per query, imgIdx 0) must hold."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle

COMMON = dict(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.data_too_large])


@st.composite
def binary_pair(draw):
    cols = draw(st.sampled_from([32, 61, 64]))
    nq, nt = draw(st.integers(0, 48)), draw(st.integers(0, 48))
    seed = draw(st.integers(0, 2**31 - 1))
    levels = draw(st.sampled_from([2, 4, 256]))  # 2 levels: heavy ties
    rng = np.random.default_rng(seed)
    Q = (rng.integers(0, levels, (nq, cols)) * (255 // (levels - 1))).astype(np.uint8)
    T = (rng.integers(0, levels, (nt, cols)) * (255 // (levels - 1))).astype(np.uint8)
    if nt >= 4 and nq >= 1 and draw(st.booleans()):  # duplicates in the train set and an exact copy of a query
        T[nt - 1] = T[0]
        T[nt // 2] = Q[0]
    return Q, T


@st.composite
def float_pair(draw):
    nq, nt = draw(st.integers(0, 40)), draw(st.integers(0, 40))
    seed = draw(st.integers(0, 2**31 - 1))
    rng = np.random.default_rng(seed)
    Q = np.floor(rng.random((nq, 128), dtype=np.float32) * 8).astype(np.float32)  # small integers: exact sums, many ties
    T = np.floor(rng.random((nt, 128), dtype=np.float32) * 8).astype(np.float32)
    if nt >= 3 and nq >= 1:
        T[nt - 1] = T[0]
        T[1] = Q[0]
    return Q, T


def _check_structure(m, nq, nt):
    assert m.dtype == oracle.DMATCH_DTYPE
    assert (m["imgIdx"] == 0).all()
    assert (np.diff(m["queryIdx"]) > 0).all()  # ascending, at most one entry per query
    assert len(m) <= nq
    if nt < 2:
        assert len(m) == 0  # no second neighbour: defined as "no matches" (the reference indexes [i][1] there)
    else:
        assert ((m["trainIdx"] >= 0) & (m["trainIdx"] < nt)).all() and ((m["queryIdx"] >= 0) & (m["queryIdx"] < nq)).all()


@settings(**COMMON)
@given(binary_pair(), st.booleans(), st.sampled_from([0.6, 0.8, 1.0]))
def test_binary_restatements_agree(pair, cross, ratio):
    Q, T = pair
    a = oracle.match_pair(Q, T, oracle.NORM_HAMMING, ratio, cross)
    b = oracle.match_pair_np(Q, T, oracle.NORM_HAMMING, ratio, cross)
    assert a.tobytes() == b.tobytes()
    _check_structure(a, len(Q), len(T))
    if not cross:
        assert a.tobytes() == oracle.match_pair_c_single(Q, T, oracle.NORM_HAMMING, ratio, False).tobytes()
    if oracle.have_cv2() and len(Q) and len(T) >= 2:
        assert a.tobytes() == oracle.match_pair_cv2(Q, T, oracle.NORM_HAMMING, ratio, cross).tobytes()
    if len(T) >= 2 and len(Q):
        dist, idx = oracle.knn2_c(Q, T, oracle.NORM_HAMMING)
        assert (dist[:, 0] <= dist[:, 1]).all()
        tie = dist[:, 0] == dist[:, 1]
        assert (idx[tie, 0] < idx[tie, 1]).all()  # lowest train index first in BOTH slots
        # a tie fails the ratio test unless both distances are 0 (ratio < 1)
        if ratio < 1.0 and not cross:
            kept = np.isin(np.arange(len(Q)), a["queryIdx"])
            assert not (kept & tie & (dist[:, 0] > 0)).any()


@settings(**COMMON)
@given(binary_pair())
def test_row_pitch_and_padding_do_not_change_the_result(pair):
    Q, T = pair
    cols = Q.shape[1]
    pitch = (cols + 15) // 16 * 16 + 16
    Qp = np.zeros((len(Q), pitch), np.uint8)
    Tp = np.zeros((len(T), pitch), np.uint8)
    Qp[:, :cols] = Q
    Tp[:, :cols] = T
    a = oracle.match_pair(Q, T, oracle.NORM_HAMMING)
    assert a.tobytes() == oracle.match_pair(Qp[:, :cols], Tp[:, :cols], oracle.NORM_HAMMING).tobytes()  # strided views
    assert a.tobytes() == oracle.match_pair(Qp, Tp, oracle.NORM_HAMMING).tobytes()                      # zero padding is neutral


@settings(**COMMON)
@given(float_pair(), st.booleans())
def test_float_integer_valued_restatements_agree_bit_for_bit(pair, cross):
    Q, T = pair
    a = oracle.match_pair(Q, T, oracle.NORM_L2, 0.8, cross)
    assert a.tobytes() == oracle.match_pair_np(Q, T, oracle.NORM_L2, 0.8, cross).tobytes()
    _check_structure(a, len(Q), len(T))
    if oracle.have_cv2() and len(Q) and len(T) >= 2:
        assert a.tobytes() == oracle.match_pair_cv2(Q, T, oracle.NORM_L2, 0.8, cross).tobytes()


def test_cross_check_is_a_subset_of_the_ratio_list():
    rng = np.random.default_rng(0)
    Q = rng.integers(0, 256, (200, 61), dtype=np.uint8)
    T = np.concatenate([Q[:80] ^ (rng.random((80, 61)) < 0.02).astype(np.uint8), rng.integers(0, 256, (150, 61), dtype=np.uint8)])
    plain = oracle.match_pair(Q, T, oracle.NORM_HAMMING, 0.8, False)
    crossed = oracle.match_pair(Q, T, oracle.NORM_HAMMING, 0.8, True)
    assert len(crossed) <= len(plain) and len(plain) >= 70
    key = lambda m: set(zip(m["queryIdx"].tolist(), m["trainIdx"].tolist()))  # noqa: E731
    assert key(crossed) <= key(plain)
