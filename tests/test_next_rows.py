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
