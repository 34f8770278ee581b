"""CPU oracle for the all-pairs descriptor-matching path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; the product package
``sfm_danpipeline_b200`` never does and has no CPU fallback.

Three restatements of ``StructFromMotion::getMatching`` (/root/reference/src/Sfm.cpp:590-608)
live here, each checked against the others and against the committed cv2 golden vectors:

* ``bf_oracle.c``   plain C (built by ``oracle/Makefile`` into ``oracle/_build/liboracle.so``),
* ``match_pair_np`` numpy, small cases,
* ``match_pair_cv2`` the OpenCV code the reference itself calls (``cv::BFMatcher::knnMatch``
  -> ``cv::batchDistance``), reached through the cv2 wheel when it is importable.

Parity pin: the reference ships no tests; the pin is cv2 4.13.0 run in the build container
(``tests/golden/make_golden.py``), see ``bf_oracle.c``'s header.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

NORM_HAMMING = 0
NORM_L2 = 1

DMATCH_DTYPE = np.dtype(
    [("queryIdx", "<i4"), ("trainIdx", "<i4"), ("imgIdx", "<i4"), ("distance", "<f4")]
)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile bf_oracle.c (gcc, no external libraries)."""
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(
        os.path.join(_HERE, "bf_oracle.c")
    ):
        subprocess.check_call(["make", "-C", _HERE, "CC=gcc"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        vp, i32, sz, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float
        for name in ("oracle_knn2_hamming", "oracle_knn2_l2"):
            getattr(L, name).argtypes = [vp, i32, sz, vp, i32, sz, i32, vp, vp]
            getattr(L, name).restype = None
        for name in ("oracle_colmin_hamming", "oracle_colmin_l2"):
            getattr(L, name).argtypes = [vp, i32, sz, vp, i32, sz, i32, vp]
            getattr(L, name).restype = None
        L.oracle_match_pair.argtypes = [vp, i32, sz, vp, i32, sz, i32, i32, f32, i32, vp]
        L.oracle_match_pair.restype = ctypes.c_int
        _lib = L
    return _lib


def _check(Q: np.ndarray, T: np.ndarray, norm: int):
    want = np.uint8 if norm == NORM_HAMMING else np.float32
    l2_on_bytes = norm == NORM_L2 and Q.dtype == np.uint8 and T.dtype == np.uint8  # see _widen
    if not l2_on_bytes and (Q.dtype != want or T.dtype != want):
        raise TypeError(f"norm {norm} needs {want} rows, got {Q.dtype}/{T.dtype}")
    if Q.ndim != 2 or T.ndim != 2 or Q.shape[1] != T.shape[1]:
        raise ValueError("descriptor sets must be 2-D with equal width")
    if (Q.size and Q.strides[1] != Q.itemsize) or (T.size and T.strides[1] != T.itemsize):
        raise ValueError("rows must be contiguous")


def _widen(Q, T, norm):
    """NORM_L2 on CV_8U rows -- what the reference's cv::BFMatcher(cv::NORM_L2) (src/Sfm.cpp:593) computes for its
    AKAZE / ORB detectors (src/Sfm.cpp:331-384).  OpenCV's batchDistL2_8u32f is sqrtf of the sum of squared byte
    differences; that sum is an integer below 2^24 for rows of up to 258 bytes, so accumulating the same integers
    in fp32 is exact in any order: widening to float32 and taking the L2 restatement is bit-identical
    (tests/test_oracle.py checks it against cv2)."""
    if norm == NORM_L2 and Q.dtype == np.uint8:
        return Q.astype(np.float32), T.astype(np.float32)
    return Q, T


# ----------------------------------------------------------------------------- C oracle
def knn2_c(Q: np.ndarray, T: np.ndarray, norm: int, threads: int = 1):
    """batchDistance(K=2): returns (dist[nq,2], idx[nq,2]); dist is int32 (Hamming) or float32."""
    _check(Q, T, norm)
    Q, T = _widen(Q, T, norm)
    L = lib()
    nq, nt, cols = Q.shape[0], T.shape[0], Q.shape[1]
    idx = np.empty((nq, 2), np.int32)
    dist = np.empty((nq, 2), np.int32 if norm == NORM_HAMMING else np.float32)
    fn = L.oracle_knn2_hamming if norm == NORM_HAMMING else L.oracle_knn2_l2

    def run(lo, hi):
        if hi > lo:
            fn(Q[lo:hi].ctypes.data, hi - lo, Q.strides[0], T.ctypes.data, nt, T.strides[0] if nt else 0,
               cols, dist[lo:hi].ctypes.data, idx[lo:hi].ctypes.data)

    if threads <= 1 or nq < 4 * threads:
        run(0, nq)
    else:
        cuts = np.linspace(0, nq, 4 * threads + 1).astype(int)
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda ab: run(*ab), zip(cuts[:-1], cuts[1:])))
    return dist, idx


def colmin_c(Q: np.ndarray, T: np.ndarray, norm: int, threads: int = 1) -> np.ndarray:
    """For each train row the lowest-index nearest query row (cross-check half)."""
    _check(Q, T, norm)
    Q, T = _widen(Q, T, norm)
    L = lib()
    nq, nt, cols = Q.shape[0], T.shape[0], Q.shape[1]
    best = np.empty(nt, np.int32)
    fn = L.oracle_colmin_hamming if norm == NORM_HAMMING else L.oracle_colmin_l2

    def run(lo, hi):
        if hi > lo:
            fn(Q.ctypes.data, nq, Q.strides[0] if nq else 0, T[lo:hi].ctypes.data, hi - lo, T.strides[0],
               cols, best[lo:hi].ctypes.data)

    if threads <= 1 or nt < 4 * threads:
        run(0, nt)
    else:
        cuts = np.linspace(0, nt, 4 * threads + 1).astype(int)
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda ab: run(*ab), zip(cuts[:-1], cuts[1:])))
    return best


def _filter(dist, idx, ratio, best_q=None) -> np.ndarray:
    """Ratio test of src/Sfm.cpp:603-607 (fp32 product, '<=') + optional mutual-NN filter."""
    d = dist.astype(np.float32)
    keep = (idx[:, 1] >= 0) & (d[:, 0] <= np.float32(ratio) * d[:, 1])
    if best_q is not None:
        safe = np.where(idx[:, 0] >= 0, idx[:, 0], 0)
        keep &= best_q[safe] == np.arange(len(idx), dtype=np.int32)
    q = np.nonzero(keep)[0].astype(np.int32)
    out = np.zeros(len(q), DMATCH_DTYPE)
    out["queryIdx"] = q
    out["trainIdx"] = idx[q, 0]
    out["distance"] = d[q, 0]
    return out


def match_pair(Q, T, norm, ratio=0.8, cross_check=False, threads: int = 1) -> np.ndarray:
    """getMatching for one ordered pair through the C oracle (threaded over query rows)."""
    _check(Q, T, norm)
    if Q.shape[0] == 0 or T.shape[0] < 2:
        return np.zeros(0, DMATCH_DTYPE)
    dist, idx = knn2_c(Q, T, norm, threads)
    best_q = colmin_c(Q, T, norm, threads) if cross_check else None
    return _filter(dist, idx, ratio, best_q)


def match_pair_c_single(Q, T, norm, ratio=0.8, cross_check=False) -> np.ndarray:
    """The single C function oracle_match_pair (no Python in the arithmetic)."""
    _check(Q, T, norm)
    Q, T = _widen(Q, T, norm)
    out = np.zeros(max(Q.shape[0], 1), DMATCH_DTYPE)
    n = lib().oracle_match_pair(Q.ctypes.data, Q.shape[0], Q.strides[0] if Q.shape[0] else 0,
                                T.ctypes.data, T.shape[0], T.strides[0] if T.shape[0] else 0,
                                Q.shape[1], norm, ratio, int(cross_check), out.ctypes.data)
    return out[:n].copy()


def all_pairs(descs, norm, ratio=0.8, cross_check=False, threads: int = 1):
    """findBestPair's q<t loop (src/Sfm.cpp:511-515): {(q,t): matches}."""
    n = len(descs)
    return {(q, t): match_pair(descs[q], descs[t], norm, ratio, cross_check, threads)
            for q in range(n - 1) for t in range(q + 1, n)}


# ------------------------------------------------------------------------- numpy oracle
_POP8 = np.array([bin(i).count("1") for i in range(256)], np.int32)


def knn2_np(Q, T, norm):
    """numpy restatement of batchDistance(K=2) with strict-'<' insertion (small cases)."""
    _check(Q, T, norm)
    Q, T = _widen(Q, T, norm)
    nq, nt = Q.shape[0], T.shape[0]
    if norm == NORM_HAMMING:
        D = _POP8[Q[:, None, :] ^ T[None, :, :]].sum(-1, dtype=np.int32) if nt else np.zeros((nq, 0), np.int32)
        big = np.iinfo(np.int32).max
    else:
        diff = Q[:, None, :] - T[None, :, :]
        D = np.sqrt((diff * diff).sum(-1, dtype=np.float32)).astype(np.float32)
        big = np.finfo(np.float32).max
    dist = np.full((nq, 2), big, D.dtype)
    idx = np.full((nq, 2), -1, np.int32)
    if nt >= 1:
        # stable argsort == ascending distance, then ascending train index
        order = np.argsort(D, axis=1, kind="stable")[:, :2]
        k = order.shape[1]
        idx[:, :k] = order
        dist[:, :k] = np.take_along_axis(D, order, 1)
    return dist, idx


def match_pair_np(Q, T, norm, ratio=0.8, cross_check=False) -> np.ndarray:
    _check(Q, T, norm)
    if Q.shape[0] == 0 or T.shape[0] < 2:
        return np.zeros(0, DMATCH_DTYPE)
    dist, idx = knn2_np(Q, T, norm)
    best_q = knn2_np(T, Q, norm)[1][:, 0] if cross_check else None
    return _filter(dist, idx, ratio, best_q)


# --------------------------------------------------------------------------- cv2 oracle
def have_cv2() -> bool:
    try:
        import cv2  # noqa: F401
        return True
    except Exception:
        return False


def knn2_cv2(Q, T, norm):
    """cv2.batchDistance(K=2): the arrays BFMatcher::knnMatch wraps into DMatches."""
    import cv2
    if norm == NORM_HAMMING:
        return cv2.batchDistance(Q, T, cv2.CV_32S, K=2, normType=cv2.NORM_HAMMING)
    return cv2.batchDistance(Q, T, cv2.CV_32F, K=2, normType=cv2.NORM_L2)


def match_pair_cv2(Q, T, norm, ratio=0.8, cross_check=False) -> np.ndarray:
    """The reference's own call sequence: BFMatcher(norm,false).knnMatch(k=2) + ratio loop."""
    import cv2
    _check(Q, T, norm)
    if Q.shape[0] == 0 or T.shape[0] < 2:
        return np.zeros(0, DMATCH_DTYPE)
    dist, idx = knn2_cv2(np.ascontiguousarray(Q), np.ascontiguousarray(T), norm)
    best_q = None
    if cross_check:
        cvn = cv2.NORM_HAMMING if norm == NORM_HAMMING else cv2.NORM_L2
        mm = cv2.BFMatcher(cvn, True).match(np.ascontiguousarray(Q), np.ascontiguousarray(T))
        best_q = np.full(T.shape[0], -1, np.int32)
        for m in mm:
            best_q[m.trainIdx] = m.queryIdx
    return _filter(dist, idx, ratio, best_q)


def knnmatch_cv2_objects(Q, T, norm):
    """Literal BFMatcher.knnMatch -> list[list[DMatch]] (used once, to pin batchDistance)."""
    import cv2
    cvn = cv2.NORM_HAMMING if norm == NORM_HAMMING else cv2.NORM_L2
    return cv2.BFMatcher(cvn, False).knnMatch(Q, T, 2)
