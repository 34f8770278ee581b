"""GPU parity of the next row of the path, descriptor extraction (getFeature, /root/reference/src/Sfm.cpp:303-392, ORB branch
:358-384): the CUDA extractor through the C ABI vs cv::ORB's output on the reference's data/temple images (committed golden)
and vs the numpy oracle on other inputs.  Bar: the same keypoints on every pyramid level and, at each of them, bit-identical
angle, Harris response, size and descriptor.  Order inside a level is not part of the contract (see sfm_features.h)."""
import os

import numpy as np
import pytest

from oracle import orb_oracle as O
from sfm_danpipeline_b200 import Matcher, NORM_HAMMING, OrbExtractor, SfmmError, extract_features

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def golden():
    z = np.load(os.path.join(HERE, "golden", "temple_orb_features.npz"))
    return z, np.concatenate([[0], np.cumsum(z["counts"])])


def as_dict(x, y, size, angle, response, octave, desc):
    return {(int(o), float(a), float(b)): (float(an), float(r), float(s), bytes(d)) for a, b, s, an, r, o, d in zip(x, y, size, angle, response, octave, desc)}


def gpu_dict(kp, desc):
    assert (kp["class_id"] == -1).all()
    return as_dict(kp["x"], kp["y"], kp["size"], kp["angle"], kp["response"], kp["octave"], desc)


def test_temple_equals_cv2_orb_golden():
    z, offs = golden()
    with OrbExtractor() as orb:
        for i in range(len(z["images"])):
            kp, desc = orb.detectAndCompute(z["images"][i])
            ref_kp, ref_d = z["keypoints"][offs[i]:offs[i + 1]], z["descriptors"][offs[i]:offs[i + 1]]
            ref = as_dict(ref_kp[:, 0], ref_kp[:, 1], ref_kp[:, 2], ref_kp[:, 3], ref_kp[:, 4], ref_kp[:, 5].astype(int), ref_d)
            got = gpu_dict(kp, desc)
            assert len(kp) == len(ref_kp) == 500
            assert set(got) == set(ref), i          # keypoint set equality, level by level
            assert got == ref, i                    # bit-exact angle / response / size / descriptor at equal keypoints
            # output order: level by level, row-major
            key = kp["octave"].astype(np.int64) << 40 | np.rint(kp["y"] / 1.2 ** kp["octave"]).astype(np.int64) << 20 | np.rint(kp["x"] / 1.2 ** kp["octave"]).astype(np.int64)
            assert (np.diff(key) > 0).all()
        assert orb.stats()["kernel_launches"] >= 51


def test_bgr_input_is_converted_like_cvtcolor():
    z, offs = golden()
    with OrbExtractor() as orb:
        a = orb.detectAndCompute(z["bgr0"])
        b = orb.detectAndCompute(z["images"][0])
        assert a[0].tobytes() == b[0].tobytes() and a[1].tobytes() == b[1].tobytes()


@pytest.mark.parametrize("shape", [(480, 640), (300, 417), (720, 1280)])
def test_other_sizes_and_strides_equal_the_oracle(shape):
    rng = np.random.default_rng(shape[0])
    z, _ = golden()
    base = z["images"][3]
    big = np.kron(base, np.ones((2, 2), np.uint8))[: max(shape[0], 1), : max(shape[1], 1)]
    img = np.ascontiguousarray(big[: shape[0], : shape[1]])
    img = (img.astype(np.int32) + rng.integers(-6, 7, img.shape)).clip(0, 255).astype(np.uint8)
    wide = np.zeros((shape[0], shape[1] + 37), np.uint8)
    wide[:, : shape[1]] = img
    with OrbExtractor() as orb:
        kp, desc = orb.detectAndCompute(wide[:, : shape[1]])  # strided rows
    okp, od = O.detect_and_compute(img)
    ref = as_dict(okp["x"], okp["y"], okp["size"], okp["angle"], okp["response"], okp["octave"], od)
    got = gpu_dict(kp, desc)
    assert set(got) == set(ref) and got == ref


def test_flat_image_ties_capacity_and_errors():
    with OrbExtractor() as orb:
        kp, desc = orb.detectAndCompute(np.full((200, 300), 77, np.uint8))  # no corners at all
        assert len(kp) == 0 and desc.shape == (0, 32)
        with pytest.raises(SfmmError) as e:
            orb.detectAndCompute(np.zeros((10, 10), np.float32))
        assert e.value.code == -1
        z, _ = golden()
        kps = np.zeros(10, orb.detectAndCompute(z["images"][0])[0].dtype)
        import ctypes as C
        n = C.c_int32()
        d = np.zeros((10, 32), np.uint8)
        rc = orb._L.sfmm_orb_detect_and_compute(orb._o, z["images"][0].ctypes.data, 480, 640, 640, 1, kps.ctypes.data, d.ctypes.data, 10, C.byref(n))
        assert rc == -5 and n.value == 500 and not d.any()  # capacity too small: size reported, nothing written


def test_extract_then_match_is_the_reference_front_end():
    """extractFeature (src/Sfm.cpp:257-298) + findBestPair's matching loop (:511-515), both on the GPU: the descriptors feed
    the matcher directly and the aligned points come from the extracted keypoints."""
    import oracle
    z, offs = golden()
    kps, descs, pts = extract_features(list(z["images"][:4]))
    with Matcher(NORM_HAMMING) as m:
        m.set_descriptors(descs)
        m.set_points(pts)
        m.match_all_pairs()
        for q in range(3):
            for t in range(q + 1, 4):
                got = m.getMatching(q, t)
                assert got.tobytes() == oracle.match_pair(descs[q], descs[t], 0).tobytes()
                left, right = m.aligned_points(q, t)
                assert (left == pts[q][got["queryIdx"]]).all() and (right == pts[t][got["trainIdx"]]).all()
