"""GPU parity, binary descriptors: the CUDA path through the C ABI vs the oracle / cv2 goldens.

Bar: bit-exact match indices and Hamming distances (north_star)."""
import numpy as np
import pytest

import oracle
from sfm_danpipeline_b200 import BINARY_AUTO, BINARY_POPC, Matcher, NORM_HAMMING, SfmmError, synth
from _golden import GoldenSet

pytestmark = pytest.mark.gpu


_ENGINE = BINARY_AUTO


@pytest.fixture(autouse=True, params=[BINARY_POPC, BINARY_AUTO], ids=["popc", "auto"])
def engine(request):
    """Every test of the generic part runs on the XOR+POPC kernel and on the default (AUTO: the tcgen05 i8
    engine for <= 512-bit descriptors, POPC above)."""
    global _ENGINE
    _ENGINE = request.param
    return request.param


def _matcher(*args, **kw):
    kw.setdefault("binary_engine", _ENGINE)
    return Matcher(NORM_HAMMING, *args, **kw)


def _expect_equal(got, exp):
    assert got.dtype == exp.dtype
    assert got.tobytes() == exp.tobytes(), (len(got), len(exp))


@pytest.mark.parametrize("name", ["temple_akaze", "temple_orb", "synth_binary"])
@pytest.mark.parametrize("cross", [False, True])
def test_all_pairs_equal_cv2_golden(name, cross):
    g = GoldenSet(name)
    with _matcher(0.8, cross) as m:
        m.set_descriptors(g.descs)
        m.match_all_pairs()
        for p, (q, t, *_r) in enumerate(g.pairs):
            got = m.getMatching(q, t)
            eq, et, ed = g.expected(p, cross)
            assert (got["queryIdx"] == eq).all() and (got["trainIdx"] == et).all(), (name, q, t)
            assert (got["distance"] == ed).all() and (got["imgIdx"] == 0).all()


@pytest.mark.parametrize("name", ["temple_akaze", "temple_orb"])
def test_raw_knn_equals_cv2_golden(name):
    g = GoldenSet(name)
    with _matcher() as m:
        m.set_descriptors(g.descs)
        for q, t, kd, ki, *_r in g.pairs[::5]:
            idx, dist = m.knn_pair(q, t)
            assert (idx == ki).all() and (dist == kd).all()


@pytest.mark.parametrize("cols", [16, 32, 61, 64, 100, 128])
def test_widths_and_ties_vs_oracle(cols):
    rng = np.random.default_rng(cols)
    # few distinct byte values => many exact distance ties, exercising lowest-index tie-breaking
    descs = [rng.integers(0, 2, (n, cols), dtype=np.uint8) * 255 for n in (700, 513, 1024, 3)]
    for cross in (False, True):
        with _matcher(0.8, cross) as m:
            m.set_descriptors(descs)
            m.match_all_pairs()
            for (q, t) in synth.all_pairs(len(descs)):
                _expect_equal(m.getMatching(q, t), oracle.match_pair(descs[q], descs[t], 0, 0.8, cross, threads=4))


def test_cfg2_shape_sample_vs_oracle():
    # configs[1]: 486-bit AKAZE-shape, 5k rows per image; a few pairs at full size
    descs = synth.binary_images(4, 5000, seed=0)
    with _matcher() as m:
        m.set_descriptors(descs)
        m.match_all_pairs()
        for (q, t) in [(0, 1), (1, 3), (2, 3)]:
            exp = oracle.match_pair(descs[q], descs[t], 0, 0.8, False, threads=8)
            _expect_equal(m.getMatching(q, t), exp)
            assert 400 < len(exp) < 900  # the generator's ~n/8 ratio-passing matches
        idx, dist = m.knn_pair(0, 1)
        d, i = oracle.knn2_c(descs[0], descs[1], 0, threads=8)
        assert (idx == i).all() and (dist == d.astype(np.float32)).all()
        ties = (d[:, 0] == d[:, 1]).mean()
        assert ties > 0.02  # the data really exercises tie-breaking


@pytest.mark.parametrize("rows", [10000, 20000])
def test_cfg3_cfg5_row_counts_vs_oracle(rows):
    """One full-size pair at configs[2] (10 000 rows) and configs[4] (20 000 rows) per engine, byte-compared with the
    oracle, raw 2-NN lists included; plus a ragged partner so that the last query tile and the last train tile are partial."""
    import os
    threads = os.cpu_count() or 8
    descs = synth.binary_images(3, [rows, rows, rows - 4321], seed=rows)
    with _matcher() as m:
        m.set_descriptors(descs)
        m.match_all_pairs()
        for (q, t) in [(0, 1), (1, 2)]:
            exp = oracle.match_pair(descs[q], descs[t], 0, 0.8, False, threads=threads)
            _expect_equal(m.getMatching(q, t), exp)
            assert rows // 16 < len(exp) < rows // 4
        idx, dist = m.knn_pair(2, 0)   # never in the q<t table: on demand, reverse direction
        d, i = oracle.knn2_c(descs[2], descs[0], 0, threads=threads)
        assert (idx == i).all() and (dist == d.astype(np.float32)).all()
    with _matcher(0.8, True) as m:     # cross-check at full size
        m.set_descriptors(descs[:2])
        _expect_equal(m.match_pair(0, 1), oracle.match_pair(descs[0], descs[1], 0, 0.8, True, threads=threads))


def test_ragged_and_degenerate_images():
    rng = np.random.default_rng(1)
    descs = [rng.integers(0, 256, (n, 61), dtype=np.uint8) for n in (0, 1, 2, 130, 1500, 0, 37)]
    for cross in (False, True):
        with _matcher(0.9, cross) as m:
            m.set_descriptors(descs)
            m.match_all_pairs()
            for (q, t) in synth.all_pairs(len(descs)):
                _expect_equal(m.getMatching(q, t), oracle.match_pair(descs[q], descs[t], 0, 0.9, cross))
                _expect_equal(m.match_pair(q, t), oracle.match_pair(descs[q], descs[t], 0, 0.9, cross))
            # reverse direction and self pairs on demand
            _expect_equal(m.match_pair(4, 3), oracle.match_pair(descs[4], descs[3], 0, 0.9, cross))
            _expect_equal(m.match_pair(4, 4), oracle.match_pair(descs[4], descs[4], 0, 0.9, cross))
            idx, dist = m.knn_pair(3, 1)  # one train row: second neighbour missing
            assert (idx[:, 1] == -1).all() and (dist[:, 1] == np.finfo(np.float32).max).all()
            assert (idx[:, 0] == 0).all()


def test_strided_rows_and_padding_are_neutral():
    rng = np.random.default_rng(2)
    wide = rng.integers(0, 256, (400, 80), dtype=np.uint8)
    a, b = wide[:200, :61], wide[200:, :61]  # row step 80, 61 used bytes
    with _matcher() as m:
        m.set_descriptors([a, b])
        got = m.match_pair(0, 1)
    _expect_equal(got, oracle.match_pair(np.ascontiguousarray(a), np.ascontiguousarray(b), 0))


def test_duplicate_rows_lowest_index_in_both_slots():
    T = np.zeros((300, 61), np.uint8)
    T[7, 0] = 1
    Q = np.zeros((2, 61), np.uint8)
    Q[1, 0] = 1
    with _matcher() as m:
        m.set_descriptors([Q, T])
        idx, dist = m.knn_pair(0, 1)
        assert idx.tolist() == [[0, 1], [7, 0]] and dist.tolist() == [[0.0, 0.0], [0.0, 1.0]]
        got = m.match_pair(0, 1)  # 0 <= 0.8*0 passes; 0 <= 0.8*1 passes
        assert got["trainIdx"].tolist() == [0, 7]


def test_batched_launches_give_the_same_table():
    descs = synth.binary_images(7, [300, 650, 1, 512, 513, 90, 1200], seed=3)
    with _matcher() as a, _matcher(pair_batch=4) as b:
        a.set_descriptors(descs)
        b.set_descriptors(descs)
        a.match_all_pairs()
        b.match_all_pairs()
        ta, tb = a.result_table(), b.result_table()
        for x, y in zip(ta, tb):
            assert x.tobytes() == y.tobytes()
        pairs, counts, offs, mat = ta
        assert pairs.tolist() == [list(p) for p in synth.all_pairs(7)]
        for i, (q, t) in enumerate(pairs):
            _expect_equal(mat[offs[i]:offs[i] + counts[i]], a.getMatching(q, t))


def test_error_codes():
    with _matcher() as m:
        with pytest.raises(SfmmError) as e:
            m.match_all_pairs()
        assert e.value.code == -4  # no descriptors yet
        with pytest.raises(SfmmError) as e:
            m.set_descriptors([np.zeros((4, 128), np.float32)])
        assert e.value.code == -1  # Hamming needs uint8
        m.set_descriptors([np.zeros((4, 61), np.uint8), np.zeros((5, 61), np.uint8)])
        with pytest.raises(SfmmError) as e:
            m.getMatching(0, 1)
        assert e.value.code == -4  # not computed yet
        with pytest.raises(SfmmError) as e:
            m.match_pair(0, 2)
        assert e.value.code == -5
        with pytest.raises(SfmmError) as e:
            m.set_descriptors([np.zeros((1 << 18, 32), np.uint8)])
        assert e.value.code == -5  # OpenCV's 2^18 row limit


def test_full_size_properties_cfg2():
    # size-independent checks at configs[1] scale: ascending queryIdx, one entry per query,
    # self-match of an image finds itself at distance 0, permutation equivariance of the train set
    descs = synth.binary_images(3, 5000, seed=9)
    perm = np.random.default_rng(0).permutation(5000)
    with _matcher() as m:
        m.set_descriptors([descs[0], descs[1], descs[1][perm]])
        m.match_all_pairs()
        a, b = m.getMatching(0, 1), m.getMatching(0, 2)
        assert (np.diff(a["queryIdx"]) > 0).all()
        assert len(a) == len(b) and (a["queryIdx"] == b["queryIdx"]).all() and (a["distance"] == b["distance"]).all()
        # same train rows up to ties broken by position
        same = perm[b["trainIdx"]] == a["trainIdx"]
        assert same.mean() > 0.95
        idx, dist = m.knn_pair(1, 1)
        assert (dist[:, 0] == 0).all()


# ----------------------------------------------------------------- opt-in tensor-core engine (tcgen05 kind::i8)
from sfm_danpipeline_b200 import BINARY_TENSOR  # noqa: E402


@pytest.mark.parametrize("name", ["temple_akaze", "temple_orb", "synth_binary"])
@pytest.mark.parametrize("cross", [False, True])
def test_tensor_engine_equals_cv2_golden(name, cross):
    g = GoldenSet(name)
    with Matcher(NORM_HAMMING, 0.8, cross, binary_engine=BINARY_TENSOR) as m:
        m.set_descriptors(g.descs)
        m.match_all_pairs()
        assert m.stats()["float_path"] == 2  # the tcgen05 kernel really ran
        for p, (q, t, kd, ki, *_r) in enumerate(g.pairs):
            got = m.getMatching(q, t)
            eq, et, ed = g.expected(p, cross)
            assert (got["queryIdx"] == eq).all() and (got["trainIdx"] == et).all(), (name, q, t)
            assert (got["distance"] == ed).all() and (got["imgIdx"] == 0).all()
        for q, t, kd, ki, *_r in g.pairs[::9]:
            idx, dist = m.knn_pair(q, t)
            assert (idx == ki).all() and (dist == kd).all()


@pytest.mark.parametrize("f4x", ["0", "1"])
@pytest.mark.parametrize("cols", [16, 29, 32, 61, 64])
def test_tensor_engine_widths_ties_ragged(cols, f4x, monkeypatch):
    # f4x = 1 forces TM_F4X where the rows have the 17 spare elements it needs (16 and 29 bytes: one K-block; 61: two), the others
    # stay on TM_F4P (32 bytes: no spare) / kind::i8 (64 bytes)
    monkeypatch.setenv("SFMM_F4X", f4x)
    rng = np.random.default_rng(cols)
    descs = [rng.integers(0, 2, (n, cols), dtype=np.uint8) * 255 for n in (700, 513, 1024, 3, 0, 129)]
    descs[4] = np.zeros((0, cols), np.uint8)
    for cross in (False, True):
        with Matcher(NORM_HAMMING, 0.8, cross, binary_engine=BINARY_TENSOR) as m, Matcher(NORM_HAMMING, 0.8, cross, binary_engine=BINARY_POPC) as ref:
            m.set_descriptors(descs)
            ref.set_descriptors(descs)
            m.match_all_pairs()
            ref.match_all_pairs()
            for a, b in zip(m.result_table(), ref.result_table()):
                assert a.tobytes() == b.tobytes()  # both engines: identical tables
            _expect_equal(m.getMatching(0, 2), oracle.match_pair(descs[0], descs[2], 0, 0.8, cross, threads=4))


def test_tensor_engine_cfg2_shape_and_limits():
    descs = synth.binary_images(4, 5000, seed=0)
    with Matcher(NORM_HAMMING, binary_engine=BINARY_TENSOR) as m:
        m.set_descriptors(descs)
        m.match_all_pairs()
        for (q, t) in [(0, 1), (2, 3)]:
            _expect_equal(m.getMatching(q, t), oracle.match_pair(descs[q], descs[t], 0, 0.8, False, threads=8))
    with Matcher(NORM_HAMMING, binary_engine=BINARY_TENSOR) as m:
        m.set_descriptors([np.zeros((10, 100), np.uint8)] * 2)  # 800 bit: beyond the engine's 512
        with pytest.raises(SfmmError) as e:
            m.match_all_pairs()
        assert e.value.code == -1


def test_incremental_match_pairs_like_addmoreviews():
    # addMoreViews / find2D3DMatches (src/Sfm.cpp:964-977, 1020-1042) ask for one new view against the done views:
    # the table grows call by call and earlier pairs stay addressable
    descs = synth.binary_images(5, [400, 380, 512, 90, 700], seed=17)
    with _matcher() as m:
        m.set_descriptors(descs)
        m.match_pairs([(0, 1)])
        first = m.getMatching(0, 1)
        m.match_pairs([(0, 2), (1, 2)])
        m.match_pairs([(2, 4), (0, 4), (1, 4)])
        assert m.getMatching(0, 1).tobytes() == first.tobytes()
        for (q, t) in [(0, 1), (0, 2), (1, 2), (2, 4), (0, 4), (1, 4)]:
            _expect_equal(m.getMatching(q, t), oracle.match_pair(descs[q], descs[t], 0))
        pairs, counts, offs, mat = m.result_table()
        assert pairs.tolist() == [[0, 1], [0, 2], [1, 2], [2, 4], [0, 4], [1, 4]]
        assert int(counts.sum()) == len(mat) and (np.diff(offs) == counts[:-1]).all()
        with pytest.raises(SfmmError):
            m.getMatching(3, 4)  # never asked
        m.clear_results()
        with pytest.raises(SfmmError):
            m.getMatching(0, 1)


# --------------------------------------------------------------------------- randomised differential
from hypothesis import HealthCheck, given, settings  # noqa: E402
from hypothesis import strategies as st  # noqa: E402


@settings(max_examples=20, deadline=None, suppress_health_check=list(HealthCheck))
@given(st.sampled_from([32, 61, 64]), st.lists(st.integers(0, 300), min_size=2, max_size=4), st.integers(0, 2**31 - 1), st.booleans(),
       st.sampled_from([2, 256]))
def test_randomised_differential_vs_oracle(cols, rows, seed, cross, levels):
    rng = np.random.default_rng(seed)
    descs = [(rng.integers(0, levels, (n, cols)) * (255 // (levels - 1))).astype(np.uint8) for n in rows]
    if rows[0] >= 1 and rows[1] >= 3:  # duplicates in a train set and an exact copy of a query row
        descs[1][-1] = descs[1][0]
        descs[1][1] = descs[0][0]
    with _matcher(0.8, cross) as m:
        m.set_descriptors(descs)
        m.match_all_pairs()
        for q, t in synth.all_pairs(len(descs)):
            _expect_equal(m.getMatching(q, t), oracle.match_pair(descs[q], descs[t], 0, 0.8, cross))


@pytest.mark.parametrize("full", ["0", "1"])
def test_cross_check_candidate_columns_and_full_reverse_agree(full, monkeypatch):
    """The tensor engines take the symmetric cross-check's column minima from "reverse" work items: by default only for the
    candidate train rows (those a ratio-passing query row selected; rows gathered through per-pair lists), with
    SFMM_CROSS_FULL=1 for every train row (round 1).  Both must equal cv2's crossCheck=True lists, ragged / empty / tiny images
    and heavy ties included."""
    monkeypatch.setenv("SFMM_CROSS_FULL", full)
    for name in ("temple_akaze", "temple_orb"):
        g = GoldenSet(name)
        with Matcher(NORM_HAMMING, 0.8, True, binary_engine=BINARY_TENSOR) as m:
            m.set_descriptors(g.descs)
            m.match_all_pairs()
            for p, (q, t, *_r) in enumerate(g.pairs):
                got = m.getMatching(q, t)
                eq, et, ed = g.expected(p, True)
                assert (got["queryIdx"] == eq).all() and (got["trainIdx"] == et).all() and (got["distance"] == ed).all(), (name, q, t)
    rng = np.random.default_rng(99)
    descs = [rng.integers(0, 2, (n, 61), dtype=np.uint8) * 255 for n in (700, 0, 513, 1, 2, 1024, 129, 5000)]  # few values: many ties
    descs[1] = np.zeros((0, 61), np.uint8)
    with Matcher(NORM_HAMMING, 0.9, True, binary_engine=BINARY_TENSOR) as m:
        m.set_descriptors(descs)
        m.match_all_pairs()
        for (q, t) in synth.all_pairs(len(descs)):
            _expect_equal(m.getMatching(q, t), oracle.match_pair(descs[q], descs[t], 0, 0.9, True, threads=4))
        _expect_equal(m.match_pair(7, 5), oracle.match_pair(descs[7], descs[5], 0, 0.9, True, threads=4))  # on demand, q > t
    sift = GoldenSet("temple_sift")
    with Matcher(1, 0.8, True) as m:  # float path (fp16 tensor kernel)
        m.set_descriptors(sift.descs)
        m.match_all_pairs()
        for p, (q, t, *_r) in enumerate(sift.pairs):
            got = m.getMatching(q, t)
            eq, et, ed = sift.expected(p, True)
            assert (got["queryIdx"] == eq).all() and (got["trainIdx"] == et).all() and (got["distance"] == ed).all(), (q, t)


# ----------------------------------------------------------------- round 2: the FP4 pipe (kind::mxf4) behind the tensor engine
@pytest.mark.parametrize("env,kind", [({}, 2), ({"SFMM_F4X": "1"}, 2), ({"SFMM_F4X": "1", "SFMM_EPI_GROUPS": "2"}, 2), ({"SFMM_NO_F4": "1"}, 1),
                                      ({"SFMM_EPI_GROUPS": "3"}, 2), ({"SFMM_F4X": "1", "SFMM_NO_SKIP": "1"}, 2)])
def test_tensor_engine_kinds_agree_with_cv2_and_oracle(env, kind, monkeypatch):
    """Descriptors below 512 bit run on the FP4 pipe: TM_F4P (packed keys, every column folded), or -- by default once an image has 7 000
    rows, forced here with SFMM_F4X=1 -- TM_F4X (key term in the MMA + threshold-skipping epilogue; AKAZE has the 17 spare elements it
    needs, ORB has none and stays on TM_F4P); or on kind::i8 (SFMM_NO_F4=1); two or three epilogue groups.  The settings are read when the context is created; every variant must reproduce the cv2 goldens, cross-check
    included, and the oracle on ragged / tied / duplicated rows."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    for name in ("temple_akaze", "temple_orb"):
        g = GoldenSet(name)
        for cross in (False, True):
            with Matcher(NORM_HAMMING, 0.8, cross) as m:
                m.set_descriptors(g.descs)
                m.match_all_pairs()
                assert m.stats()["float_path"] == 2 and m.stats()["tensor_kind"] == kind
                for p, (q, t, *_r) in enumerate(g.pairs):
                    got = m.getMatching(q, t)
                    eq, et, ed = g.expected(p, cross)
                    assert (got["queryIdx"] == eq).all() and (got["trainIdx"] == et).all() and (got["distance"] == ed).all(), (name, q, t, cross)
    rng = np.random.default_rng(7)
    base = rng.integers(0, 256, (3000, 61), dtype=np.uint8)
    base[:, 60] &= 0x3F  # AKAZE's two pad bits
    dup = np.concatenate([base[:50]] * 3 + [base[50:1500]])  # triplicated rows: ties in both slots
    extremes = np.concatenate([np.zeros((2, 61), np.uint8), np.full((2, 61), 255, np.uint8), base[:200]])
    extremes[2:4, 60] = 0x3F
    descs = [base, dup, base[::-1].copy(), extremes, base[:1], np.zeros((0, 61), np.uint8), base[:129]]
    for cross in (False, True):
        with Matcher(NORM_HAMMING, 0.8, cross) as m:
            m.set_descriptors(descs)
            m.match_all_pairs()
            for (q, t) in synth.all_pairs(len(descs)):
                _expect_equal(m.getMatching(q, t), oracle.match_pair(descs[q], descs[t], 0, 0.8, cross, threads=8))


@pytest.mark.parametrize("order", ["descending", "ascending"])
def test_threshold_skipping_is_order_independent(order, monkeypatch):
    """The skipping epilogue's cost depends on the order the train rows arrive in, its result must not: train rows sorted by their
    distance to one query row -- nearest last (every block takes the slow path) or nearest first (none does after the first tile) --
    and many exact ties."""
    monkeypatch.setenv("SFMM_F4X", "1")
    rng = np.random.default_rng(11)
    q = rng.integers(0, 256, (300, 61), dtype=np.uint8)
    t = rng.integers(0, 256, (6000, 61), dtype=np.uint8)
    t[::7] = t[3]  # ties
    q[:, 60] &= 0x3F
    t[:, 60] &= 0x3F
    d = np.unpackbits(t ^ q[0], axis=1).sum(1)
    idx = np.argsort(d, kind="stable")
    t = t[idx[::-1] if order == "descending" else idx].copy()
    for cross in (False, True):
        with Matcher(NORM_HAMMING, 0.8, cross) as m:
            m.set_descriptors([q, t])
            m.match_all_pairs()
            assert m.stats()["tensor_kind"] == 2
            _expect_equal(m.getMatching(0, 1), oracle.match_pair(q, t, 0, 0.8, cross, threads=8))


def test_f4x_key_term_covers_every_popcount(monkeypatch):
    """TM_F4X writes 512 - popc(t) into the spare elements of every train row as E2M1 digits (binary_unpack4x_kernel): every popcount
    0..486 must come out right, with query rows of every popcount as well (popc(q) completes the distance in the epilogue)."""
    monkeypatch.setenv("SFMM_F4X", "1")
    rng = np.random.default_rng(3)
    rows = []
    for k in range(487):
        bits = np.zeros(488, np.uint8)
        bits[rng.choice(486, k, replace=False)] = 1
        rows.append(np.packbits(bits, bitorder="little"))
    t = np.stack(rows)
    q = t[rng.permutation(487)].copy()
    q[::3] ^= rng.integers(0, 256, q[::3].shape, dtype=np.uint8)
    q[:, 60] &= 0x3F
    for cross in (False, True):
        with Matcher(NORM_HAMMING, 0.95, cross) as m:
            m.set_descriptors([q, t, t[::-1].copy()])
            m.match_all_pairs()
            assert m.stats()["tensor_kind"] == 2
            for (a, b) in synth.all_pairs(3):
                d = [q, t, t[::-1].copy()]
                _expect_equal(m.getMatching(a, b), oracle.match_pair(d[a], d[b], 0, 0.95, cross, threads=4))
                idx, dist = m.knn_pair(a, b)
                odist, oidx = oracle.knn2_c(d[a], d[b], 0)
                assert (idx == oidx).all() and (dist == odist).all()
