"""GPU tests of the rows SURVEY.md section 8(f) marks "next": aligned keypoints gathered on the device
(AlignedPointsFromMatch, /root/reference/src/Sfm.cpp:694-711) and the persisted match table."""
import numpy as np
import pytest

import oracle
from sfm_danpipeline_b200 import Matcher, NORM_HAMMING, NORM_L2, SfmmError, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["binary", "float"])
def test_aligned_points_equal_the_reference_gather(kind):
    if kind == "binary":
        descs, norm = synth.binary_images(5, [900, 0, 1300, 257, 2100], seed=31), NORM_HAMMING
        descs[1] = np.zeros((0, 61), np.uint8)
    else:
        descs, norm = synth.float_images(4, [700, 300, 129, 1100], seed=32), NORM_L2
    rng = np.random.default_rng(0)
    pts = [rng.random((len(d), 2)) * [640.0, 480.0] for d in descs]  # imagesPts2D
    with Matcher(norm, 0.8, kind == "binary") as m:
        m.set_descriptors(descs)
        m.set_points(pts)
        m.match_all_pairs()
        for (q, t) in synth.all_pairs(len(descs)):
            mm = m.getMatching(q, t)
            exp = oracle.match_pair(descs[q], descs[t], 0 if kind == "binary" else 1, 0.8, kind == "binary")
            assert mm.tobytes() == exp.tobytes()
            left, right = m.aligned_points(q, t)
            # AlignedPoints: alignedL[i] = queryImg[matches[i].queryIdx], alignedR[i] = trainImg[matches[i].trainIdx]
            assert (left == pts[q][mm["queryIdx"]]).all() and (right == pts[t][mm["trainIdx"]]).all()


def test_points_must_be_set_before_matching():
    descs = synth.binary_images(2, 300, seed=1)
    with Matcher(NORM_HAMMING) as m:
        m.set_descriptors(descs)
        m.match_all_pairs()
        with pytest.raises(SfmmError) as e:
            m.aligned_points(0, 1)
        assert e.value.code == -4
        with pytest.raises(SfmmError):
            m.set_points([np.zeros((5, 2))] * 2)  # one point per descriptor row


def test_persisted_table_round_trip(tmp_path):
    descs = synth.binary_images(6, [800, 40, 0, 1500, 513, 2], seed=33)
    descs[2] = np.zeros((0, 61), np.uint8)
    path = str(tmp_path / "matches.sfmm")
    with Matcher(NORM_HAMMING) as m:
        m.set_descriptors(descs)
        m.match_all_pairs()
        ref = m.result_table()
        m.save_table(path)
    with Matcher(NORM_HAMMING) as m2:
        m2.set_descriptors(descs)
        m2.load_table(path)  # no kernel runs
        got = m2.result_table()
        for a, b in zip(ref, got):
            assert a.tobytes() == b.tobytes()
        for (q, t) in synth.all_pairs(len(descs)):
            assert m2.getMatching(q, t).tobytes() == oracle.match_pair(descs[q], descs[t], 0).tobytes()
        assert m2.stats()["pairs_matched"] == 0
    # a table of another descriptor set, or a damaged file, is refused
    with Matcher(NORM_HAMMING) as m3:
        m3.set_descriptors(descs[:5])
        with pytest.raises(SfmmError) as e:
            m3.load_table(path)
        assert e.value.code == -1
        m3.set_descriptors(descs)
        raw = open(path, "rb").read()
        open(path, "wb").write(raw[: len(raw) // 2])
        with pytest.raises(SfmmError) as e:
            m3.load_table(path)
        assert e.value.code == -1


def test_persisted_table_is_bound_to_its_filter_settings_and_survives_corrupt_headers(tmp_path):
    """a table computed with another ratio / cross-check must not be served, and a header with absurd
    sizes must come back as an error code (nothing may throw across the C ABI)."""
    import struct
    descs = synth.binary_images(3, [300, 280, 150], seed=34)
    path = str(tmp_path / "t.sfmm")
    with Matcher(NORM_HAMMING, 0.8, False) as m:
        m.set_descriptors(descs)
        m.match_all_pairs()
        m.save_table(path)
    for ratio, cross in [(0.7, False), (0.8, True)]:
        with Matcher(NORM_HAMMING, ratio, cross) as m:
            m.set_descriptors(descs)
            with pytest.raises(SfmmError) as e:
                m.load_table(path)
            assert e.value.code == -1 and "ratio" in str(e.value)
    raw = bytearray(open(path, "rb").read())
    with Matcher(NORM_HAMMING, 0.8, False) as m:
        m.set_descriptors(descs)
        m.load_table(path)
        with pytest.raises(SfmmError) as e:  # the file holds no aligned points
            m.aligned_points(0, 1)
        assert e.value.code == -4
        for n_pairs, n_matches in [(2 ** 62, 0), (3, 2 ** 61), (-1, 0), (3, 7)]:
            bad = bytearray(raw)
            bad[32:48] = struct.pack("<qq", n_pairs, n_matches)  # header: magic[8] 4xi32 f32 i32 | n_pairs n_matches
            open(path, "wb").write(bad)
            with pytest.raises(SfmmError) as e:
                m.load_table(path)
            assert e.value.code == -1
        assert len(m.getMatching(0, 1)) == len(oracle.match_pair(descs[0], descs[1], 0))  # the loaded table is still served


def test_group_matcher_single_process_all_gpus():
    """sfmm_group_*: every visible GPU from one process -- one H2D, NCCL broadcast inside the library, sharded pairs."""
    import torch
    from sfm_danpipeline_b200 import GroupMatcher
    n_dev = torch.cuda.device_count()
    rows = [900, 0, 1300, 257, 64, 2100, 31]
    descs = synth.binary_images(len(rows), rows, seed=51)
    descs[1] = np.zeros((0, 61), np.uint8)
    for cross in (False, True):
        with GroupMatcher(n_dev, NORM_HAMMING, 0.8, cross) as g:
            assert g.size == n_dev
            for _ in range(2):  # twice: buffers reused, re-broadcast
                g.set_descriptors(descs)
                g.match_all_pairs()
            ts = g.transfer_stats()
            blob = sum((r + 3) // 4 * 4 for r in rows) * 64  # device layout: 64-byte rows, images aligned to four rows
            assert ts["h2d_bytes"] == sum(rows) * 61          # uploaded once, as the caller's tightly packed 61-byte rows
            assert ts["nccl_bytes"] == blob * (n_dev - 1)      # broadcast to every other device
            for (q, t) in synth.all_pairs(len(rows)):
                assert g.getMatching(q, t).tobytes() == oracle.match_pair(descs[q], descs[t], 0, 0.8, cross).tobytes(), (q, t, cross)
            with pytest.raises(SfmmError) as e:
                g.getMatching(3, 1)  # q>t: not in findBestPair's table
            assert e.value.code == -4
            assert sum(g.member_stats(i)["pairs_matched"] for i in range(n_dev)) == 2 * len(synth.all_pairs(len(rows)))
    fd = synth.float_images(4, [300, 200, 150, 90], seed=52)
    with GroupMatcher(n_dev) as g:  # the reference's constants: NORM_L2, 0.8, no cross-check
        g.set_descriptors(fd)
        g.match_pairs([(0, 1), (2, 3), (1, 3)])
        for (q, t) in [(0, 1), (2, 3), (1, 3)]:
            assert g.getMatching(q, t).tobytes() == oracle.match_pair(fd[q], fd[t], 1).tobytes()
