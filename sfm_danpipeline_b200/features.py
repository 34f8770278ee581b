"""Host-side mirror of the reference's feature-extraction entry point, ORB branch, on top of the C ABI (include/sfm_features.h).

Reference interface (C++): ``void StructFromMotion::getFeature(const cv::Mat& image, const int& numImage)``
(/root/reference/src/Sfm.cpp:303-392) with ``detector == 3``: ``cv::ORB::create(500, 1.2f, 8, 31, 0, 2, HARRIS_SCORE, 31, 20)``
+ ``detectAndCompute`` (:360-373), filling ``imagesKeypoints`` / ``imagesDescriptors`` / ``imagesPts2D`` (:380-382).
All arithmetic runs in the CUDA library -- there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import SfmmError

#: numpy view of SfmKeyPoint == cv::KeyPoint
KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4"),
                           ("class_id", "<i4")])


class OrbExtractor:
    """One ``cv::ORB`` with the reference's parameters on one B200."""

    def __init__(self, device: int = 0):
        self._L = _lib.load()
        self._o = C.c_void_p()
        rc = self._L.sfmm_orb_create(int(device), C.byref(self._o))
        if rc != 0:
            raise SfmmError(rc, self._L.sfmm_orb_last_error(None).decode())

    def close(self):
        if getattr(self, "_o", None) and self._o.value:
            self._L.sfmm_orb_destroy(self._o)
            self._o = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def detectAndCompute(self, image: np.ndarray, capacity: int = 1024):
        """(keypoints as KEYPOINT_DTYPE records, descriptors (n, 32) uint8) of a gray (h, w) or BGR (h, w, 3) uint8 image."""
        if image.dtype != np.uint8 or image.ndim not in (2, 3) or (image.ndim == 3 and image.shape[2] != 3):
            raise SfmmError(_lib.SFMM_EINVAL, "image must be (h, w) or (h, w, 3) uint8")
        ch = 1 if image.ndim == 2 else 3
        if image.strides[-1] != 1 or (ch == 3 and image.strides[1] != 3):
            image = np.ascontiguousarray(image)
        kps = np.zeros(capacity, KEYPOINT_DTYPE)
        desc = np.zeros((capacity, 32), np.uint8)
        n = C.c_int32()
        rc = self._L.sfmm_orb_detect_and_compute(self._o, image.ctypes.data, image.shape[0], image.shape[1], image.strides[0], ch,
                                                 kps.ctypes.data, desc.ctypes.data, capacity, C.byref(n))
        if rc == _lib.SFMM_ERANGE and n.value > capacity:
            return self.detectAndCompute(image, n.value)
        if rc != 0:
            raise SfmmError(rc, self._L.sfmm_orb_last_error(self._o).decode())
        return kps[: n.value].copy(), desc[: n.value].copy()

    def stats(self) -> dict:
        k, ms = C.c_int64(), C.c_double()
        self._L.sfmm_orb_stats(self._o, C.byref(k), C.byref(ms))
        return {"kernel_launches": k.value, "last_ms": ms.value}


def extract_features(images, device: int = 0):
    """extractFeature's loop (src/Sfm.cpp:257-298) for the ORB detector: returns (imagesKeypoints, imagesDescriptors, imagesPts2D)."""
    kps_all, desc_all, pts_all = [], [], []
    with OrbExtractor(device) as orb:
        for img in images:
            k, d = orb.detectAndCompute(img)
            kps_all.append(k)
            desc_all.append(d)
            pts_all.append(np.stack([k["x"], k["y"]], 1).astype(np.float64))  # keypointstoPoints, src/Sfm.cpp:397-403
    return kps_all, desc_all, pts_all
