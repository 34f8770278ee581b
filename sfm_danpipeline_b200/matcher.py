"""Host-side mirror of the reference's matching interface on top of the C ABI.

Reference interface (C++): ``void StructFromMotion::getMatching(const int& idx_query,
const int& idx_train, Matching* goodMatches)`` (/root/reference/include/Sfm.h:89,
src/Sfm.cpp:590-608) over ``std::vector<cv::Mat> imagesDescriptors`` (include/Sfm.h:29), driven
for all q<t by ``findBestPair`` (src/Sfm.cpp:511-515).  This class keeps the names and argument
meaning; match lists come back as numpy records with cv::DMatch's fields.  All arithmetic runs
in the CUDA library -- there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np

from . import _lib
from ._lib import DMATCH_DTYPE, NORM_HAMMING, NORM_L2, SfmmError


class Matcher:
    """One ``cv::BFMatcher(norm, false)`` + ratio loop, for every image pair, on one B200.

    norm         ``NORM_L2`` (the reference's hard-wired choice, src/Sfm.cpp:593) for float32
                 descriptors or ``NORM_HAMMING`` for uint8 (AKAZE / ORB) descriptors
    ratio        ``NN_MATCH_RATIO`` (include/Sfm.h:60)
    cross_check  north-star extra stage, off by default like the reference
    """

    def __init__(self, norm: int = NORM_L2, ratio: float = 0.8, cross_check: bool = False, device: int = 0,
                 float_mode: int = _lib.FLOAT_AUTO, pair_batch: int = 0, binary_engine: int = _lib.BINARY_AUTO):
        self._L = _lib.load()
        cfg = _lib.SfmmConfig()
        self._L.sfmm_default_config(C.byref(cfg))
        cfg.device, cfg.norm, cfg.ratio = int(device), int(norm), float(ratio)
        cfg.cross_check, cfg.float_mode, cfg.pair_batch = int(bool(cross_check)), int(float_mode), int(pair_batch)
        cfg.binary_engine = int(binary_engine)
        self._ctx = C.c_void_p()
        rc = self._L.sfmm_create(C.byref(cfg), C.byref(self._ctx))
        if rc != 0:
            raise SfmmError(rc, self._L.sfmm_last_error(None).decode())
        self.norm, self.device = int(norm), int(device)
        self.rows: list[int] = []
        self.cols = 0
        self.elem_u8 = self.norm == NORM_HAMMING

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) and self._ctx.value:
            self._L.sfmm_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise SfmmError(rc, self._L.sfmm_last_error(self._ctx).decode())

    # -- descriptors ----------------------------------------------------------------------
    def set_descriptors(self, descriptors: Sequence[np.ndarray]):
        """imagesDescriptors: one (rows_i, cols) uint8 / float32 array per image (rows may be strided)."""
        ptrs, rows, steps, cols, elem, keep = _marshal_descriptors(descriptors, self.norm)
        self._check(self._L.sfmm_set_descriptors(self._ctx, len(keep), ptrs, rows, cols, steps, elem))
        self.rows, self.cols, self.elem_u8 = [d.shape[0] for d in keep], cols, elem == _lib.U8

    def reserve_descriptors(self, rows: Sequence[int], cols: int, elem_u8: bool | None = None):
        """Allocate the device layout only (non-root ranks, before the blob broadcast)."""
        n = len(rows)
        r = (C.c_int32 * max(n, 1))(*[int(x) for x in rows])
        if elem_u8 is None:
            elem_u8 = self.norm == NORM_HAMMING
        self._check(self._L.sfmm_set_descriptors(self._ctx, n, None, r, int(cols), None, _lib.U8 if elem_u8 else _lib.F32))
        self.rows, self.cols, self.elem_u8 = [int(x) for x in rows], int(cols), bool(elem_u8)

    def descriptor_blob(self) -> tuple[int, int]:
        """(device pointer, bytes) of the packed descriptor blob."""
        p, b = C.c_void_p(), C.c_size_t()
        self._check(self._L.sfmm_descriptor_blob(self._ctx, C.byref(p), C.byref(b)))
        return int(p.value or 0), int(b.value)

    # -- matching -------------------------------------------------------------------------
    def match_all_pairs(self):
        """findBestPair's q<t loop, once; afterwards getMatching is a look-up."""
        self._check(self._L.sfmm_match_all_pairs(self._ctx))

    def match_pairs(self, pairs):
        qt = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))
        self._check(self._L.sfmm_match_pairs(self._ctx, qt.ctypes.data, len(qt)))

    def match_pairs_device(self, pairs, d_counts_ptr: int, d_matches_ptr: int, capacity: int) -> int:
        """Match into caller-owned DEVICE buffers (see sfmm_match_pairs_device); returns #records."""
        qt = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))
        n = C.c_int64()
        self._check(self._L.sfmm_match_pairs_device(self._ctx, qt.ctypes.data, len(qt), d_counts_ptr, d_matches_ptr,
                                                    int(capacity), C.byref(n)))
        return int(n.value)

    def getMatching(self, idx_query: int, idx_train: int) -> np.ndarray:
        """The reference's entry point: the ratio-filtered 1-NN list of (idx_query, idx_train)."""
        p, n = C.c_void_p(), C.c_int32()
        rc = self._L.sfmm_get_pair(self._ctx, int(idx_query), int(idx_train), C.byref(p), C.byref(n))
        if rc != 0:  # look-ups do not set the library's error text
            raise SfmmError(rc, {_lib.SFMM_ESTATE: "pair has not been matched (match_all_pairs / match_pairs first)",
                                 _lib.SFMM_ERANGE: "image index out of range"}.get(rc, "getMatching failed"))
        if n.value == 0:
            return np.zeros(0, DMATCH_DTYPE)
        buf = (C.c_char * (n.value * DMATCH_DTYPE.itemsize)).from_address(p.value)
        return np.frombuffer(buf, DMATCH_DTYPE).copy()

    get_matching = getMatching

    def match_pair(self, idx_query: int, idx_train: int) -> np.ndarray:
        """getMatching computed on demand (no table)."""
        cap = max(self.rows[idx_query], 1)
        out = np.zeros(cap, DMATCH_DTYPE)
        n = C.c_int32()
        self._check(self._L.sfmm_match_pair(self._ctx, int(idx_query), int(idx_train), out.ctypes.data, cap, C.byref(n)))
        return out[: n.value].copy()

    def knn_pair(self, idx_query: int, idx_train: int):
        """knnMatch(k=2) as arrays: (train_idx[nq,2] int32, distance[nq,2] float32)."""
        nq = self.rows[idx_query]
        idx = np.empty((nq, 2), np.int32)
        dist = np.empty((nq, 2), np.float32)
        self._check(self._L.sfmm_knn_pair(self._ctx, int(idx_query), int(idx_train), idx.ctypes.data, dist.ctypes.data))
        return idx, dist

    def result_table(self, copy: bool = True):
        """(pairs[n,2], counts[n], offsets[n], matches[total]) of everything matched so far.

        copy=False returns read-only views of the library's host table (valid until the next
        match / set_descriptors / clear_results / close)."""
        n, m = C.c_int64(), C.c_int64()
        qt, cnt, off, mat = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self._L.sfmm_result_table(self._ctx, C.byref(n), C.byref(qt), C.byref(cnt), C.byref(off),
                                              C.byref(mat), C.byref(m)))

        def arr(ptr, count, dt):
            if count == 0:
                return np.zeros(0, dt)
            buf = (C.c_char * (count * np.dtype(dt).itemsize)).from_address(ptr.value)
            a = np.frombuffer(buf, dt)
            if copy:
                return a.copy()
            a.flags.writeable = False
            return a

        return (arr(qt, 2 * n.value, np.int32).reshape(-1, 2), arr(cnt, n.value, np.int32),
                arr(off, n.value, np.int64), arr(mat, m.value, DMATCH_DTYPE))

    # -- next rows of the path ------------------------------------------
    def set_points(self, points):
        """imagesPts2D: one (rows_i, 2) float64 array per image (cv::Point2d).  Call after
        set_descriptors; matching then also gathers AlignedPointsFromMatch's lists on the GPU."""
        keep = [np.ascontiguousarray(p, np.float64).reshape(-1, 2) for p in points]
        if [len(p) for p in keep] != list(self.rows):
            raise SfmmError(_lib.SFMM_EINVAL, "one point per descriptor row is required")
        ptrs = (C.c_void_p * max(len(keep), 1))(*[p.ctypes.data if len(p) else None for p in keep])
        self._check(self._L.sfmm_set_points(self._ctx, len(keep), ptrs))

    def aligned_points(self, idx_query: int, idx_train: int):
        """(alignedL, alignedR): (count, 2) float64 arrays in match order (src/Sfm.cpp:694-711)."""
        l, r, n = C.c_void_p(), C.c_void_p(), C.c_int32()
        rc = self._L.sfmm_get_pair_points(self._ctx, int(idx_query), int(idx_train), C.byref(l), C.byref(r), C.byref(n))
        if rc != 0:
            raise SfmmError(rc, "pair not matched with points (set_points before matching; not restored by load_table)")
        if n.value == 0:
            return np.zeros((0, 2)), np.zeros((0, 2))

        def arr(p):
            buf = (C.c_char * (n.value * 16)).from_address(p.value)
            return np.frombuffer(buf, np.float64).reshape(-1, 2).copy()

        return arr(l), arr(r)

    def save_table(self, path: str):
        self._check(self._L.sfmm_save_table(self._ctx, str(path).encode()))

    def load_table(self, path: str):
        self._check(self._L.sfmm_load_table(self._ctx, str(path).encode()))

    def clear_results(self):
        self._check(self._L.sfmm_clear_results(self._ctx))

    def share_table(self, shm_prefix: str | None):
        """Keep the match table in a POSIX shared-memory segment other processes of this host can map (see sfm_match.h)."""
        self._check(self._L.sfmm_share_table(self._ctx, shm_prefix.encode() if shm_prefix else None))

    def shared_table_info(self) -> tuple[str, int]:
        """(segment name, records in it) of the shared table; the name is '' until something has been matched."""
        buf, n = C.create_string_buffer(256), C.c_int64()
        self._check(self._L.sfmm_shared_table_info(self._ctx, buf, 256, C.byref(n)))
        return buf.value.decode(), int(n.value)

    def stats(self) -> dict:
        s = _lib.SfmmStats()
        self._check(self._L.sfmm_get_stats(self._ctx, C.byref(s)))
        return {k: getattr(s, k) for k, _ in s._fields_}


def _marshal_descriptors(descriptors, norm):
    """(ptrs, rows, steps, cols, elem) ctypes arguments of sfmm_set_descriptors + the arrays that must stay alive."""
    n = len(descriptors)
    want = np.uint8 if norm == NORM_HAMMING else (np.uint8 if n and descriptors[0].dtype == np.uint8 else np.float32)
    cols = descriptors[0].shape[1] if n else 1
    keep = []
    for d in descriptors:
        if d.ndim != 2 or d.dtype != want or d.shape[1] != cols:
            raise SfmmError(_lib.SFMM_EINVAL, f"every descriptor set must be (rows, {cols}) {want.__name__}")
        if d.shape[0] and d.strides[1] != d.itemsize:
            d = np.ascontiguousarray(d)
        keep.append(d)
    ptrs = (C.c_void_p * max(n, 1))(*[d.ctypes.data if d.shape[0] else None for d in keep])
    rows = (C.c_int32 * max(n, 1))(*[d.shape[0] for d in keep])
    steps = (C.c_size_t * max(n, 1))(*[d.strides[0] if d.shape[0] > 1 else d.shape[1] * d.itemsize for d in keep])
    return ptrs, rows, steps, cols, (_lib.U8 if want is np.uint8 else _lib.F32), keep


class GroupMatcher:
    """Every B200 of the box from ONE process (sfmm_group_*): descriptors uploaded once and broadcast with NCCL
    (single-process ncclCommInitAll inside the library), pairs dealt to the devices, getMatching as a look-up.
    The shape a patched single-process iTree3DMap would use; torch is not involved."""

    def __init__(self, n_devices: int, norm: int = NORM_L2, ratio: float = 0.8, cross_check: bool = False,
                 devices: Sequence[int] | None = None, float_mode: int = _lib.FLOAT_AUTO, binary_engine: int = _lib.BINARY_AUTO):
        self._L = _lib.load()
        cfg = _lib.SfmmConfig()
        self._L.sfmm_default_config(C.byref(cfg))
        cfg.norm, cfg.ratio, cfg.cross_check = int(norm), float(ratio), int(bool(cross_check))
        cfg.float_mode, cfg.binary_engine = int(float_mode), int(binary_engine)
        devs = (C.c_int32 * n_devices)(*[int(d) for d in devices]) if devices is not None else None
        self._g = C.c_void_p()
        rc = self._L.sfmm_group_create(C.byref(cfg), int(n_devices), devs, C.byref(self._g))
        if rc != 0:
            raise SfmmError(rc, self._L.sfmm_group_last_error(None).decode())
        self.norm = int(norm)
        self.rows: list[int] = []

    def close(self):
        if getattr(self, "_g", None) and self._g.value:
            self._L.sfmm_group_destroy(self._g)
            self._g = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc: int):
        if rc != 0:
            raise SfmmError(rc, self._L.sfmm_group_last_error(self._g).decode())

    @property
    def size(self) -> int:
        return int(self._L.sfmm_group_size(self._g))

    def set_descriptors(self, descriptors: Sequence[np.ndarray]):
        ptrs, rows, steps, cols, elem, keep = _marshal_descriptors(descriptors, self.norm)
        self._check(self._L.sfmm_group_set_descriptors(self._g, len(keep), ptrs, rows, cols, steps, elem))
        self.rows = [d.shape[0] for d in keep]

    def match_all_pairs(self):
        self._check(self._L.sfmm_group_match_all_pairs(self._g))

    def match_pairs(self, pairs):
        qt = np.ascontiguousarray(np.asarray(pairs, np.int32).reshape(-1, 2))
        self._check(self._L.sfmm_group_match_pairs(self._g, qt.ctypes.data, len(qt)))

    def getMatching(self, idx_query: int, idx_train: int) -> np.ndarray:
        p, n = C.c_void_p(), C.c_int32()
        rc = self._L.sfmm_group_get_pair(self._g, int(idx_query), int(idx_train), C.byref(p), C.byref(n))
        if rc != 0:
            raise SfmmError(rc, "pair has not been matched" if rc == _lib.SFMM_ESTATE else "image index out of range")
        if n.value == 0:
            return np.zeros(0, DMATCH_DTYPE)
        buf = (C.c_char * (n.value * DMATCH_DTYPE.itemsize)).from_address(p.value)
        return np.frombuffer(buf, DMATCH_DTYPE).copy()

    def transfer_stats(self) -> dict:
        a, b = C.c_int64(), C.c_int64()
        self._check(self._L.sfmm_group_transfer_stats(self._g, C.byref(a), C.byref(b)))
        return {"h2d_bytes": a.value, "nccl_bytes": b.value}

    def member_stats(self, i: int) -> dict:
        s = _lib.SfmmStats()
        ctx = self._L.sfmm_group_context(self._g, int(i))
        rc = self._L.sfmm_get_stats(ctx, C.byref(s))
        if rc != 0:
            raise SfmmError(rc, "no such member")
        return {k: getattr(s, k) for k, _ in s._fields_}
