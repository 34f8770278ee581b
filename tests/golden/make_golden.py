"""Generate the committed golden vectors under tests/golden/ -- run in the BUILD container only.

It needs two things that do not travel to the GPU box: /root/reference/data/temple (the
reference's only fixture, src/Sfm.cpp:118-198 loads it) and the cv2 wheel.  Everything it
writes is derived from the OpenCV code the reference itself calls:

    cv::BFMatcher(norm,false).knnMatch(q, t, knn, 2)        /root/reference/src/Sfm.cpp:593,599
    ratio loop  d1 <= 0.8f * d2                              /root/reference/src/Sfm.cpp:603-607
    all pairs q<t                                            /root/reference/src/Sfm.cpp:511-515

Detector parameters are the reference's (src/Sfm.cpp:309-313 SIFT, :333-339 AKAZE, :360-368 ORB);
images are read like imagesLOAD does (sorted by name, BGR->gray, no resize because 480x640
fails `rows>480 && cols>640`, src/Sfm.cpp:153).

Usage:  python tests/golden/make_golden.py        (writes *.npz next to this file)
"""
import glob
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from sfm_danpipeline_b200 import synth  # noqa: E402

TEMPLE = "/root/reference/data/temple"
RATIO = np.float32(0.8)


def knn_literal(Q, T, cvnorm):
    """The literal reference call: BFMatcher.knnMatch -> arrays."""
    knn = cv2.BFMatcher(cvnorm, False).knnMatch(Q, T, 2)
    dist = np.array([[m.distance for m in row] for row in knn], np.float32).reshape(len(knn), -1)
    idx = np.array([[m.trainIdx for m in row] for row in knn], np.int32).reshape(len(knn), -1)
    img = np.array([[m.imgIdx for m in row] for row in knn], np.int32)
    qi = np.array([[m.queryIdx for m in row] for row in knn], np.int32)
    assert (img == 0).all() and (qi == np.arange(len(knn))[:, None]).all()
    return dist, idx


def pair_record(Q, T, cvnorm):
    dist, idx = knn_literal(Q, T, cvnorm)
    # batchDistance must be the same thing; checked on every pair we store
    bd, bi = cv2.batchDistance(Q, T, cv2.CV_32S if cvnorm == cv2.NORM_HAMMING else cv2.CV_32F,
                               K=2, normType=cvnorm)
    assert (bi == idx).all() and (bd.astype(np.float32) == dist).all()
    keep = dist[:, 0] <= RATIO * dist[:, 1]
    mutual = cv2.BFMatcher(cvnorm, True).match(Q, T)
    mset = {(m.queryIdx, m.trainIdx) for m in mutual}
    q = np.nonzero(keep)[0]
    xkeep = np.array([(int(i), int(idx[i, 0])) in mset for i in q], bool)
    return dist, idx, q.astype(np.int32), xkeep


def build_set(descs, cvnorm):
    """All-pairs record for a list of descriptor sets."""
    rows = np.array([d.shape[0] for d in descs], np.int32)
    knn_d, knn_i, mq, mx, mcount = [], [], [], [], []
    n = len(descs)
    for a in range(n - 1):
        for b in range(a + 1, n):
            dist, idx, q, xkeep = pair_record(descs[a], descs[b], cvnorm)
            knn_d.append(dist); knn_i.append(idx); mq.append(q); mx.append(xkeep); mcount.append(len(q))
    return dict(rows=rows, desc=np.concatenate(descs, 0),
                knn_dist=np.concatenate(knn_d, 0), knn_idx=np.concatenate(knn_i, 0),
                match_q=np.concatenate(mq), match_cross=np.concatenate(mx),
                match_count=np.array(mcount, np.int32))


def temple_gray():
    files = sorted(glob.glob(os.path.join(TEMPLE, "*.png")))
    assert len(files) == 10, files
    return [cv2.cvtColor(cv2.imread(f), cv2.COLOR_BGR2GRAY) for f in files]


def main_l2_on_bytes(imgs):
    """The reference's LITERAL call for its binary detectors: cv::BFMatcher(cv::NORM_L2) is hard-wired at
    src/Sfm.cpp:593 whatever `detector` is, so AKAZE (detector 2) and ORB (detector 3) rows are matched with L2
    over the bytes (OpenCV's batchDistL2_8u32f)."""
    orb = cv2.ORB_create(500, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 20)
    d = [orb.detectAndCompute(g, None)[1] for g in imgs]
    rec = build_set(d, cv2.NORM_L2)
    np.savez_compressed(os.path.join(HERE, "temple_orb_l2.npz"), **rec)
    print("orb/L2 matches", int(rec["match_count"].sum()), "cross", int(rec["match_cross"].sum()))
    akaze = cv2.AKAZE_create(cv2.AKAZE_DESCRIPTOR_MLDB, 0, 3, 0.001, 4, 4, cv2.KAZE_DIFF_PM_G2)
    d = [akaze.detectAndCompute(g, None)[1] for g in imgs[:4]]  # four images, six pairs: keeps the file small
    rec = build_set(d, cv2.NORM_L2)
    np.savez_compressed(os.path.join(HERE, "temple_akaze_l2.npz"), **rec)
    print("akaze/L2 (4 images) matches", int(rec["match_count"].sum()), "cross", int(rec["match_cross"].sum()))


def main_orb_features(imgs):
    """Golden for the next row of the path, descriptor extraction (getFeature, src/Sfm.cpp:303-392), ORB branch (:358-384):
    the gray images themselves (the GPU box has no /root/reference) and what cv::ORB::detectAndCompute returns for them."""
    orb = cv2.ORB_create(500, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 20)
    rec = {"images": np.stack(imgs)}
    counts, kp_all, d_all = [], [], []
    for g in imgs:
        kps, desc = orb.detectAndCompute(g, None)
        counts.append(len(kps))
        kp_all.append(np.array([[k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave] for k in kps], np.float32))
        d_all.append(desc)
    rec.update(counts=np.array(counts, np.int32), keypoints=np.concatenate(kp_all), descriptors=np.concatenate(d_all))
    # colour -> gray the way imread + cvtColor(BGR2GRAY) does it, for the optional device-side conversion: first image only
    files = sorted(glob.glob(os.path.join(TEMPLE, "*.png")))
    rec["bgr0"] = cv2.imread(files[0])
    np.savez_compressed(os.path.join(HERE, "temple_orb_features.npz"), **rec)
    print("orb features", counts)


def main():
    imgs = temple_gray()
    if "--only-l2-on-bytes" in sys.argv:
        return main_l2_on_bytes(imgs)
    if "--only-orb-features" in sys.argv:
        return main_orb_features(imgs)
    main_l2_on_bytes(imgs)
    main_orb_features(imgs)

    akaze = cv2.AKAZE_create(cv2.AKAZE_DESCRIPTOR_MLDB, 0, 3, 0.001, 4, 4, cv2.KAZE_DIFF_PM_G2)
    d = [akaze.detectAndCompute(g, None)[1] for g in imgs]
    assert all(x.dtype == np.uint8 and x.shape[1] == 61 for x in d)
    rec = build_set(d, cv2.NORM_HAMMING)
    np.savez_compressed(os.path.join(HERE, "temple_akaze.npz"), **rec)
    print("akaze rows", rec["rows"].tolist(), "matches", int(rec["match_count"].sum()),
          "cross", int(rec["match_cross"].sum()))

    orb = cv2.ORB_create(500, 1.2, 8, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 20)
    d = [orb.detectAndCompute(g, None)[1] for g in imgs]
    assert all(x.dtype == np.uint8 and x.shape[1] == 32 for x in d)
    rec = build_set(d, cv2.NORM_HAMMING)
    np.savez_compressed(os.path.join(HERE, "temple_orb.npz"), **rec)
    print("orb rows", rec["rows"].tolist(), "matches", int(rec["match_count"].sum()),
          "cross", int(rec["match_cross"].sum()))

    sift = cv2.SIFT_create(0, 3, 0.04, 10, 1.6)
    d = [sift.detectAndCompute(g, None)[1] for g in imgs]
    assert all(x.dtype == np.float32 and x.shape[1] == 128 for x in d)
    assert all((x == np.floor(x)).all() and x.min() >= 0 and x.max() <= 255 for x in d)
    rec = build_set(d, cv2.NORM_L2)
    rec["desc"] = rec["desc"].astype(np.uint8)  # integer valued 0..255: store compactly, exact
    np.savez_compressed(os.path.join(HERE, "temple_sift.npz"), **rec)
    print("sift rows", rec["rows"].tolist(), "matches", int(rec["match_count"].sum()),
          "cross", int(rec["match_cross"].sum()))

    # synthetic: small seeded sets straight from the package generators
    d = synth.binary_images(4, [300, 257, 1, 130], seed=7)
    rec = build_set([x for x in d if x.shape[0] >= 2], cv2.NORM_HAMMING)
    np.savez_compressed(os.path.join(HERE, "synth_binary.npz"), **rec)
    print("synth binary matches", rec["match_count"].tolist())

    d = synth.float_images(3, [200, 150, 97], seed=11, integer=False)
    rec = build_set(d, cv2.NORM_L2)
    np.savez_compressed(os.path.join(HERE, "synth_float.npz"), **rec)
    print("synth float matches", rec["match_count"].tolist())


if __name__ == "__main__":
    main()
