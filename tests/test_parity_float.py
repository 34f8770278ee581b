"""GPU parity, float descriptors (SIFT shape): CUDA path vs oracle / cv2 goldens.

Bar (north_star): L2 distances within 1e-4 relative; indices may differ only where the top-2
gap is below that tolerance; bit-exact on integer-valued (real SIFT) data."""
import numpy as np
import pytest

import oracle
from sfm_danpipeline_b200 import FLOAT_AUTO, FLOAT_EXACT, FLOAT_TENSOR, Matcher, NORM_L2, SfmmError, synth
from _golden import GoldenSet

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _check_knn_within_tolerance(idx, dist, ki, kd):
    np.testing.assert_allclose(dist, kd, rtol=RTOL)
    bad = idx != ki
    if bad.any():
        # an index may only differ where the competing distances are closer than the tolerance
        r, c = np.nonzero(bad)
        gap = np.abs(dist[r, c] - kd[r, c]) / np.maximum(kd[r, c], 1e-30)
        assert (gap <= RTOL).all()


@pytest.mark.parametrize("cross", [False, True])
def test_temple_sift_bit_exact(cross):
    g = GoldenSet("temple_sift")  # integer-valued 0..255 floats: every partial sum is exact in fp32
    with Matcher(NORM_L2, 0.8, cross, float_mode=FLOAT_EXACT) as m:
        m.set_descriptors(g.descs)
        m.match_all_pairs()
        for p, (q, t, *_r) in enumerate(g.pairs):
            got = m.getMatching(q, t)
            eq, et, ed = g.expected(p, cross)
            assert (got["queryIdx"] == eq).all() and (got["trainIdx"] == et).all(), (q, t)
            assert (got["distance"] == ed).all()


def test_temple_sift_raw_knn_bit_exact():
    g = GoldenSet("temple_sift")
    with Matcher(NORM_L2, float_mode=FLOAT_EXACT) as m:
        m.set_descriptors(g.descs)
        for q, t, kd, ki, *_r in g.pairs[::6]:
            idx, dist = m.knn_pair(q, t)
            assert (idx == ki).all() and (dist == kd).all()


def test_non_integer_floats_within_tolerance():
    g = GoldenSet("synth_float")
    with Matcher(NORM_L2, float_mode=FLOAT_EXACT) as m:
        m.set_descriptors(g.descs)
        m.match_all_pairs()
        for p, (q, t, kd, ki, *_r) in enumerate(g.pairs):
            idx, dist = m.knn_pair(q, t)
            _check_knn_within_tolerance(idx, dist, ki, kd)
            got = m.getMatching(q, t)
            eq, et, ed = g.expected(p)
            if len(got) == len(eq) and (got["queryIdx"] == eq).all():
                np.testing.assert_allclose(got["distance"], ed, rtol=RTOL)
            else:  # a ratio decision may flip only when d1 ~ 0.8*d2 within tolerance
                flip = np.setxor1d(got["queryIdx"], eq)
                margin = np.abs(kd[flip, 0] - np.float32(0.8) * kd[flip, 1]) / kd[flip, 1]
                assert (margin <= 2 * RTOL).all()


def test_cfg4_shape_sample_vs_oracle():
    descs = synth.float_images(3, [8000, 8000, 1000], seed=0)  # configs[3] shape: 8k x 128 f32, integer valued
    with Matcher(NORM_L2, float_mode=FLOAT_EXACT) as m:
        m.set_descriptors(descs)
        m.match_all_pairs()
        for (q, t) in [(0, 1), (2, 1)]:
            exp = oracle.match_pair(descs[q], descs[t], 1, 0.8, False, threads=8)
            got = m.getMatching(q, t) if q < t else m.match_pair(q, t)
            assert got.tobytes() == exp.tobytes()


def test_ragged_float_and_widths():
    rng = np.random.default_rng(5)
    for cols in (64, 128, 36):
        descs = [np.floor(rng.random((n, cols), dtype=np.float32) * 200).astype(np.float32) for n in (0, 1, 2, 65, 300)]
        for cross in (False, True):
            with Matcher(NORM_L2, 0.85, cross, float_mode=FLOAT_EXACT) as m:
                m.set_descriptors(descs)
                m.match_all_pairs()
                for (q, t) in synth.all_pairs(len(descs)):
                    exp = oracle.match_pair(descs[q], descs[t], 1, 0.85, cross)
                    assert m.getMatching(q, t).tobytes() == exp.tobytes(), (cols, q, t, cross)


# ----------------------------------------------------------------- tensor-core (tcgen05 TF32) mode
@pytest.mark.parametrize("mode", [FLOAT_TENSOR, FLOAT_AUTO])
def test_tensor_mode_temple_sift_bit_exact(mode):
    g = GoldenSet("temple_sift")
    with Matcher(NORM_L2, 0.8, False, float_mode=mode) as m:
        m.set_descriptors(g.descs)
        m.match_all_pairs()
        assert m.stats()["float_path"] == FLOAT_TENSOR  # real SIFT output is TF32-exact: AUTO must pick the tensor path
        for p, (q, t, kd, ki, *_r) in enumerate(g.pairs):
            got = m.getMatching(q, t)
            eq, et, ed = g.expected(p)
            assert (got["queryIdx"] == eq).all() and (got["trainIdx"] == et).all(), (q, t)
            assert (got["distance"] == ed).all()
        for q, t, kd, ki, *_r in g.pairs[::7]:
            idx, dist = m.knn_pair(q, t)
            assert (idx == ki).all() and (dist == kd).all()


def test_tensor_mode_cfg4_shape_vs_oracle_and_exact_kernel():
    descs = synth.float_images(3, [8000, 8000, 777], seed=0)  # configs[3] shape
    with Matcher(NORM_L2, float_mode=FLOAT_TENSOR) as mt, Matcher(NORM_L2, float_mode=FLOAT_EXACT) as mx:
        mt.set_descriptors(descs)
        mx.set_descriptors(descs)
        mt.match_all_pairs()
        mx.match_all_pairs()
        for a, b in zip(mt.result_table(), mx.result_table()):
            assert a.tobytes() == b.tobytes()  # the two kernels agree bit for bit on every pair
        exp = oracle.match_pair(descs[0], descs[1], 1, 0.8, False, threads=8)
        assert mt.getMatching(0, 1).tobytes() == exp.tobytes()
        for (q, t) in [(0, 1), (2, 0), (1, 2)]:  # raw 2-NN lists incl. the partial last tile and a short train set
            ia, da = mt.knn_pair(q, t)
            ib, db = mx.knn_pair(q, t)
            assert (ia == ib).all() and (da == db).all()


def test_tensor_mode_ties_duplicates_and_ragged():
    rng = np.random.default_rng(8)
    base = np.floor(rng.random((300, 128), dtype=np.float32) * 60).astype(np.float32)
    T = np.concatenate([base, base[:50], base[100:130]])  # duplicated train rows => exact distance ties
    Q = np.concatenate([base[:200] + (rng.random((200, 128)) < 0.02), base[:3]]).astype(np.float32)
    sets = [Q, T, base[:1], base[:2], np.zeros((0, 128), np.float32), base[:129]]
    with Matcher(NORM_L2, 0.9, False, float_mode=FLOAT_TENSOR) as m:
        m.set_descriptors(sets)
        m.match_all_pairs()
        for (q, t) in synth.all_pairs(len(sets)):
            exp = oracle.match_pair(sets[q], sets[t], 1, 0.9, False)
            assert m.getMatching(q, t).tobytes() == exp.tobytes(), (q, t)
        idx, dist = m.knn_pair(0, 1)
        d, i = oracle.knn2_c(Q, T, 1)
        assert (idx == i).all() and (dist == d).all()
        assert (d[:, 0] == d[:, 1]).sum() > 20  # ties really occur, lowest index must win in both slots


@pytest.mark.parametrize("cols", [32, 64, 96])
def test_tensor_mode_other_widths(cols):
    rng = np.random.default_rng(cols)
    descs = [np.floor(rng.random((n, cols), dtype=np.float32) * 100).astype(np.float32) for n in (400, 333, 130)]
    with Matcher(NORM_L2, float_mode=FLOAT_TENSOR) as m:
        m.set_descriptors(descs)
        m.match_all_pairs()
        for (q, t) in synth.all_pairs(3):
            assert m.getMatching(q, t).tobytes() == oracle.match_pair(descs[q], descs[t], 1).tobytes()


# ----------------------------------------------------------------- arbitrary floats: TF32 ranking + exact refinement
def _tables_equal(a, b):
    for x, y in zip(a.result_table(), b.result_table()):
        assert x.tobytes() == y.tobytes()


def test_arbitrary_floats_use_rank_and_refine_and_equal_the_exact_kernel():
    g = GoldenSet("synth_float")  # non-integer values: not TF32-exact
    for mode in (FLOAT_AUTO, FLOAT_TENSOR):
        with Matcher(NORM_L2, float_mode=mode) as m, Matcher(NORM_L2, float_mode=FLOAT_EXACT) as mx:
            m.set_descriptors(g.descs)
            mx.set_descriptors(g.descs)
            m.match_all_pairs()
            mx.match_all_pairs()
            assert m.stats()["float_path"] == 3 and mx.stats()["float_path"] == FLOAT_EXACT
            _tables_equal(m, mx)  # the refinement uses the exact kernel's arithmetic: bit-identical
            for p, (q, t, kd, ki, *_r) in enumerate(g.pairs):
                idx, dist = m.knn_pair(q, t)
                _check_knn_within_tolerance(idx, dist, ki, kd)  # and within 1e-4 of OpenCV
    # cross-check on arbitrary floats: ranking + refinement forward, the candidate columns' minima from the exact kernel (rows
    # gathered through the candidate lists) -- bit-identical to running everything on the exact kernel
    with Matcher(NORM_L2, 0.8, True, float_mode=FLOAT_AUTO) as m, Matcher(NORM_L2, 0.8, True, float_mode=FLOAT_EXACT) as mx:
        m.set_descriptors(g.descs)
        mx.set_descriptors(g.descs)
        m.match_all_pairs()
        mx.match_all_pairs()
        assert m.stats()["float_path"] == 3 and mx.stats()["float_path"] == FLOAT_EXACT
        _tables_equal(m, mx)
        assert sum(len(m.getMatching(q, t)) for q, t in synth.all_pairs(len(g.descs))) > 10


def test_arbitrary_floats_cross_check_at_cfg4_size_equals_the_exact_kernel():
    rng = np.random.default_rng(8)
    a = synth.float_images(3, [8000, 4100, 129], seed=16, integer=False)
    a[1][:5] = a[1][7]                       # identical train rows: lowest index in both directions
    a[0][11] = a[1][7]
    a.append(np.zeros((0, 128), np.float32))  # an image without descriptors
    a.append((rng.random((2, 128)) * 100).astype(np.float32))
    with Matcher(NORM_L2, 0.8, True) as m, Matcher(NORM_L2, 0.8, True, float_mode=FLOAT_EXACT) as mx:
        m.set_descriptors(a)
        mx.set_descriptors(a)
        m.match_all_pairs()
        mx.match_all_pairs()
        assert m.stats()["float_path"] == 3
        _tables_equal(m, mx)
        got, exp = m.getMatching(0, 1), oracle.match_pair(a[0], a[1], 1, 0.8, True, threads=8)
        assert abs(len(got) - len(exp)) <= 2  # vs OpenCV: decisions may flip only inside the 1e-4 tolerance
        both = np.intersect1d(got["queryIdx"], exp["queryIdx"])
        assert len(both) >= len(exp) - 2


def test_rank_and_refine_at_cfg4_size_near_duplicates_and_ties():
    rng = np.random.default_rng(3)
    a = synth.float_images(2, [8000, 4100], seed=6, integer=False)  # SIFT-scale non-integer values
    near = (a[0][:1500] + rng.normal(0, 1e-3, (1500, 128))).astype(np.float32)  # near-duplicates: d ~ 1e-2 vs |x| ~ 500
    dup = np.repeat(a[1][:3], 40, axis=0)  # 120 identical rows: candidate lists overflow -> exact rescan of the row
    sets = [a[0], np.concatenate([a[1], near, dup]), near[:130], a[1][:1]]
    with Matcher(NORM_L2, float_mode=FLOAT_TENSOR) as m, Matcher(NORM_L2, float_mode=FLOAT_EXACT) as mx:
        m.set_descriptors(sets)
        mx.set_descriptors(sets)
        m.match_all_pairs()
        mx.match_all_pairs()
        assert m.stats()["float_path"] == 3
        _tables_equal(m, mx)
        for (q, t) in [(0, 1), (2, 1), (1, 0), (0, 3)]:
            ia, da = m.knn_pair(q, t)
            ib, db = mx.knn_pair(q, t)
            assert (ia == ib).all() and (da == db).all(), (q, t)
        exp = oracle.match_pair(sets[2], sets[1], 1, 0.8, False, threads=8)
        got = m.match_pair(2, 1)
        assert (got["queryIdx"] == exp["queryIdx"]).all() and (got["trainIdx"] == exp["trainIdx"]).all()
        np.testing.assert_allclose(got["distance"], exp["distance"], rtol=RTOL)


def test_rank_and_refine_small_scale_values_and_negative_entries():
    rng = np.random.default_rng(9)
    descs = [rng.normal(0, 0.1, (n, 64)).astype(np.float32) for n in (700, 650, 129)]  # SURF-like: signed, |x| ~ 1
    with Matcher(NORM_L2, 0.9, float_mode=FLOAT_AUTO) as m, Matcher(NORM_L2, 0.9, float_mode=FLOAT_EXACT) as mx:
        m.set_descriptors(descs)
        mx.set_descriptors(descs)
        m.match_all_pairs()
        mx.match_all_pairs()
        assert m.stats()["float_path"] == 3
        _tables_equal(m, mx)


@pytest.mark.parametrize("scale", [1e3, 1e5])
def test_rank_and_refine_with_one_dominant_row(scale):
    """One row of the set dwarfs every other norm: the ranking pass's key-table offset (the set's max |x|^2) then sits far
    above the pair's own distances, and its fp32 rounding has to be part of the error bound.  scale 1e5 also leaves
    fp16's range, so the passes run on TF32 operands.  Result: still the exact kernel's table, bit for bit."""
    rng = np.random.default_rng(11)
    descs = [(rng.random((n, 128), dtype=np.float32) * 3).astype(np.float32) for n in (400, 390, 260)]
    descs[2][7] *= scale  # |x|^2 ~ 4e8 (fp16 range) / 4e12 (beyond it)
    descs[0][:50] = descs[1][:50] + rng.normal(0, 0.01, (50, 128)).astype(np.float32)  # near duplicates: tiny distances
    with Matcher(NORM_L2, 0.8, float_mode=FLOAT_AUTO) as m, Matcher(NORM_L2, 0.8, float_mode=FLOAT_EXACT) as mx:
        m.set_descriptors(descs)
        mx.set_descriptors(descs)
        m.match_all_pairs()
        mx.match_all_pairs()
        assert m.stats()["float_path"] == 3
        _tables_equal(m, mx)


def test_tensor_mode_cross_check_temple_and_ragged():
    g = GoldenSet("temple_sift")
    with Matcher(NORM_L2, 0.8, True, float_mode=FLOAT_TENSOR) as m:
        m.set_descriptors(g.descs)
        m.match_all_pairs()
        assert m.stats()["float_path"] == FLOAT_TENSOR
        for p, (q, t, *_r) in enumerate(g.pairs):
            got = m.getMatching(q, t)
            eq, et, ed = g.expected(p, True)
            assert (got["queryIdx"] == eq).all() and (got["trainIdx"] == et).all(), (q, t)
            assert (got["distance"] == ed).all()
    rng = np.random.default_rng(12)
    base = np.floor(rng.random((700, 128), dtype=np.float32) * 50).astype(np.float32)
    sets = [base[:300], np.concatenate([base[100:400], base[:40]]), base[650:], base[:1], base[5:135]]
    with Matcher(NORM_L2, 0.95, True, float_mode=FLOAT_TENSOR) as m:
        m.set_descriptors(sets)
        m.match_all_pairs()
        for (q, t) in synth.all_pairs(len(sets)):
            exp = oracle.match_pair(sets[q], sets[t], 1, 0.95, True)
            assert m.getMatching(q, t).tobytes() == exp.tobytes(), (q, t)
        assert m.match_pair(1, 0).tobytes() == oracle.match_pair(sets[1], sets[0], 1, 0.95, True).tobytes()


def test_tensor_mode_cross_check_cfg4_shape_equals_exact_kernel():
    descs = synth.float_images(3, [8000, 4100, 130], seed=5)
    with Matcher(NORM_L2, 0.8, True, float_mode=FLOAT_TENSOR) as mt, Matcher(NORM_L2, 0.8, True, float_mode=FLOAT_EXACT) as mx:
        mt.set_descriptors(descs)
        mx.set_descriptors(descs)
        mt.match_all_pairs()
        mx.match_all_pairs()
        for a, b in zip(mt.result_table(), mx.result_table()):
            assert a.tobytes() == b.tobytes()


# --------------------------------------------------------------------------- randomised differential
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@settings(max_examples=15, deadline=None, suppress_health_check=list(HealthCheck))
@given(st.sampled_from([64, 96, 128]), st.lists(st.integers(0, 300), min_size=2, max_size=3), st.integers(0, 2**31 - 1), st.booleans(),
       st.sampled_from([FLOAT_AUTO, FLOAT_EXACT]))
def test_randomised_differential_integer_valued_vs_oracle(cols, rows, seed, cross, mode):
    """Integer-valued data (the SIFT regime): every kernel -- fp16 tensor (64, 128), TF32 tensor (96), exact fp32 -- is bit-exact."""
    rng = np.random.default_rng(seed)
    descs = [np.floor(rng.random((n, cols), dtype=np.float32) * 9).astype(np.float32) for n in rows]
    if rows[0] >= 1 and rows[1] >= 3:
        descs[1][-1] = descs[1][0]
        descs[1][1] = descs[0][0]
    with Matcher(NORM_L2, 0.8, cross, float_mode=mode) as m:
        m.set_descriptors(descs)
        m.match_all_pairs()
        for q, t in synth.all_pairs(len(descs)):
            got, exp = m.getMatching(q, t), oracle.match_pair(descs[q], descs[t], 1, 0.8, cross)
            assert got.tobytes() == exp.tobytes(), (cols, rows, q, t)


# ----------------------------------------------------------------- NORM_L2 over CV_8U rows: the reference's literal call
# cv::BFMatcher(cv::NORM_L2) is hard-wired at src/Sfm.cpp:593, so its AKAZE (detector 2) and ORB (detector 3)
# descriptors are matched with L2 over the bytes.  The library widens the bytes to fp32 on upload (exact) and the
# float kernels take over; distances are sqrtf of exact integers, so the bar is bit-exact.
@pytest.mark.parametrize("name", ["temple_orb_l2", "temple_akaze_l2"])
@pytest.mark.parametrize("cross", [False, True])
@pytest.mark.parametrize("mode", [FLOAT_AUTO, FLOAT_EXACT])
def test_l2_over_bytes_equals_cv2_golden(name, cross, mode):
    g = GoldenSet(name)
    assert g.descs[0].dtype == np.uint8
    with Matcher(NORM_L2, 0.8, cross, float_mode=mode) as m:
        m.set_descriptors(g.descs)
        m.match_all_pairs()
        for p, (q, t, kd, ki, *_r) in enumerate(g.pairs):
            got = m.getMatching(q, t)
            eq, et, ed = g.expected(p, cross)
            assert (got["queryIdx"] == eq).all() and (got["trainIdx"] == et).all(), (name, q, t)
            assert (got["distance"] == ed).all() and (got["imgIdx"] == 0).all()
        for q, t, kd, ki, *_r in g.pairs[::7]:
            idx, dist = m.knn_pair(q, t)
            assert (idx == ki).all() and (dist == kd).all()


def test_l2_over_bytes_random_widths_vs_oracle():
    rng = np.random.default_rng(77)
    for cols in (16, 32, 61, 64, 100):
        descs = [rng.integers(0, 256, (n, cols), dtype=np.uint8) for n in (300, 0, 129, 2, 700)]
        descs[4][5] = descs[4][9]          # duplicate train rows: lowest index first in both slots
        descs[0][3] = descs[4][5]
        with Matcher(NORM_L2, 0.9, False) as m:
            m.set_descriptors(descs)
            m.match_all_pairs()
            for (q, t) in synth.all_pairs(len(descs)):
                exp = oracle.match_pair(descs[q], descs[t], 1, 0.9, False)
                assert m.getMatching(q, t).tobytes() == exp.tobytes(), (cols, q, t)
    with Matcher(0) as m:  # NORM_HAMMING still refuses float rows, like cv::BFMatcher
        with pytest.raises(SfmmError) as e:
            m.set_descriptors([np.zeros((4, 128), np.float32)])
        assert e.value.code == -1


@pytest.mark.parametrize("groups,no_kx", [("2", "0"), ("4", "0"), ("2", "1"), ("4", "1")])
def test_tensor_kernel_epilogue_group_variants_agree(groups, no_kx, monkeypatch):
    """The TMEM-A float kernel runs two (default) or four epilogue groups (SFMM_EPI_GROUPS) and takes the key's train-side term
    from the tensor core (TM_F16X, default) or from a table in the epilogue (SFMM_NO_KX=1, round 1's TM_F16_EXACT); the settings
    are read when the context is created.  Every variant must reproduce the cv2 goldens and the oracle at the cfg-4 row count,
    ragged tails included."""
    monkeypatch.setenv("SFMM_EPI_GROUPS", groups)
    monkeypatch.setenv("SFMM_NO_KX", no_kx)
    g = GoldenSet("temple_sift")
    with Matcher(NORM_L2, 0.8, False) as m:
        m.set_descriptors(g.descs)
        m.match_all_pairs()
        assert m.stats()["float_path"] == FLOAT_TENSOR
        for p, (q, t, *_r) in enumerate(g.pairs):
            got = m.getMatching(q, t)
            eq, et, ed = g.expected(p)
            assert (got["queryIdx"] == eq).all() and (got["trainIdx"] == et).all() and (got["distance"] == ed).all(), (q, t)
    descs = synth.float_images(3, [8000, 7777, 130], seed=5)
    for cross in (False, True):
        with Matcher(NORM_L2, 0.8, cross) as m:
            m.set_descriptors(descs)
            m.match_all_pairs()
            for (q, t) in synth.all_pairs(3):
                assert m.getMatching(q, t).tobytes() == oracle.match_pair(descs[q], descs[t], 1, 0.8, cross, threads=8).tobytes(), (q, t, cross)


@pytest.mark.parametrize("env", [{}, {"SFMM_NO_SKIP": "1"}, {"SFMM_EPI_GROUPS": "3"}, {"SFMM_EPI_GROUPS": "4"}])
def test_threshold_skipping_epilogue_variants_agree(env, monkeypatch):
    """TM_F16X's epilogue skips the blocks of columns that cannot enter a row's top-2 any more (default; two, three or four epilogue
    groups sharing their thresholds) or folds every column (SFMM_NO_SKIP=1).  Cost depends on the arrival order of the train rows,
    results must not: random order, nearest-last and nearest-first orders, exact ties and duplicated rows."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    g = GoldenSet("temple_sift")
    for cross in (False, True):
        with Matcher(NORM_L2, 0.8, cross) as m:
            m.set_descriptors(g.descs)
            m.match_all_pairs()
            assert m.stats()["float_path"] == FLOAT_TENSOR
            for p, (q, t, *_r) in enumerate(g.pairs):
                got = m.getMatching(q, t)
                eq, et, ed = g.expected(p, cross)
                assert (got["queryIdx"] == eq).all() and (got["trainIdx"] == et).all() and (got["distance"] == ed).all(), (q, t, cross)
    a, b = synth.float_images(2, [700, 6000], seed=9)
    b[::5] = b[2]  # ties
    d = ((b - a[0]) ** 2).sum(1)
    idx = np.argsort(d, kind="stable")
    sets = [b, b[idx].copy(), b[idx[::-1]].copy()]
    for cross in (False, True):
        with Matcher(NORM_L2, 0.8, cross) as m:
            m.set_descriptors([a] + sets)
            m.match_pairs([(0, 1), (0, 2), (0, 3), (3, 0)])
            for (q, t) in [(0, 1), (0, 2), (0, 3), (3, 0)]:
                descs = [a] + sets
                assert m.getMatching(q, t).tobytes() == oracle.match_pair(descs[q], descs[t], 1, 0.8, cross, threads=8).tobytes(), (q, t, cross)
