"""CPU tests: pin oracle/orb_oracle.py (numpy restatement of the reference's ORB branch, src/Sfm.cpp:358-384) to the cv2 golden
(tests/golden/temple_orb_features.npz: cv::ORB::detectAndCompute on the reference's data/temple images) and, when cv2 is
importable, every stage to the OpenCV function it restates."""
import os

import numpy as np
import pytest

from oracle import orb_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    z = np.load(os.path.join(HERE, "golden", "temple_orb_features.npz"))
    offs = np.concatenate([[0], np.cumsum(z["counts"])])
    return z, offs


def as_sets(kp, desc):
    """{(octave, x, y): (angle, response, size, descriptor bytes)} -- order inside a level is not part of the contract."""
    return {(int(k[5]), float(k[0]), float(k[1])): (float(k[3]), float(k[4]), float(k[2]), bytes(d)) for k, d in zip(kp, desc)}


@pytest.mark.parametrize("i", [0, 4, 9])
def test_oracle_reproduces_cv2_orb_on_temple(i):
    z, offs = golden()
    kp, d = O.detect_and_compute(z["images"][i])
    mine = as_sets(np.stack([kp[f] for f in ("x", "y", "size", "angle", "response")] + [kp["octave"].astype(np.float32)], 1), d)
    ref = as_sets(z["keypoints"][offs[i]:offs[i + 1]], z["descriptors"][offs[i]:offs[i + 1]])
    assert set(mine) == set(ref)          # the same keypoints on every pyramid level
    assert mine == ref                    # and bit-identical angle, Harris response, size and descriptor at each of them


def test_reference_parameters_and_level_quota():
    assert O.features_per_level() == [109, 90, 75, 63, 52, 44, 36, 31] and sum(O.features_per_level()) == 500
    assert O.level_sizes(640, 480) == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]
    assert O.pattern().shape == (256, 4) and O.pattern()[0].tolist() == [8, -3, 9, 5] and O.pattern()[-1].tolist() == [-1, -6, 0, -11]
    assert [float(x).hex() for x in O.GAUSS_7_S2[:4]] == ["0x1.1f5f620000000p-4", "0x1.0c70fc0000000p-3", "0x1.8694720000000p-3", "0x1.ba95c00000000p-3"]
    assert O.umax_table()[:16].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]


def test_stages_against_opencv_when_available():
    cv2 = pytest.importorskip("cv2")
    z, _ = golden()
    img = z["images"][2]
    levels = O.pyramid(img)
    prev = img
    for (w, h), lev in list(zip(O.level_sizes(640, 480), levels))[1:]:
        prev = cv2.resize(prev, (w, h), interpolation=cv2.INTER_LINEAR_EXACT)
        assert (prev == lev).all()
    fd = cv2.FastFeatureDetector_create(20, True)
    for lev in (levels[0], levels[5]):
        ref = {(int(k.pt[0]), int(k.pt[1])): k.response for k in fd.detect(lev, None)}
        xs, ys, sc = O.fast_keypoints(lev)
        assert ref == {(int(x), int(y)): float(s) for x, y, s in zip(xs, ys, sc)}
    k = cv2.getGaussianKernel(7, 2, cv2.CV_32F).ravel()
    assert (k == O.GAUSS_7_S2).all()
    for lev in levels[::3]:  # ORB's in-place sub-matrix blur takes the separable fp32 filter, not the fixed-point GaussianBlur path
        assert (cv2.sepFilter2D(lev, -1, k, k, borderType=cv2.BORDER_REFLECT_101) == O.gaussian_blur_7x7(lev)).all()
    rng = np.random.default_rng(0)
    for y, x in rng.integers(-200000, 200000, (3000, 2)):
        assert O.fast_atan2(y, x) == np.float32(cv2.fastAtan2(float(y), float(x)))
    assert (O.bgr_to_gray(z["bgr0"]) == cv2.cvtColor(z["bgr0"], cv2.COLOR_BGR2GRAY)).all()
    assert (O.bgr_to_gray(z["bgr0"]) == z["images"][0]).all()
