"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what
include/sfm_match.h declares, and refuses to run without a GPU (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from sfm_danpipeline_b200 import _lib, Matcher, SfmmError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    import torch
    return torch.cuda.is_available()


def test_library_is_built_and_loads():
    L = _lib.load()
    assert b"sm_100a" in L.sfmm_version()


def test_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "sfm_match.h")).read()
    declared = set(re.findall(r"SFMM_API\s+[\w\s\*]+?\b(sfmm_\w+)\s*\(", hdr))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = _lib.load()
    for name in declared:
        assert getattr(L, name) is not None


def test_feature_exports_match_header():
    hdr = open(os.path.join(ROOT, "include", "sfm_features.h")).read()
    declared = set(re.findall(r"SFMM_API\s+[\w\s\*]+?\b(sfmm_\w+)\s*\(", hdr))
    assert declared == set(_lib.FEATURE_EXPORTS), declared ^ set(_lib.FEATURE_EXPORTS)
    L = _lib.load()
    for name in declared:
        assert getattr(L, name) is not None
    for cite in ("src/Sfm.cpp:303-392", "src/Sfm.cpp:360-371", "src/Sfm.cpp:373", "include/Utilities.h:26"):
        assert cite in hdr
    from sfm_danpipeline_b200 import KEYPOINT_DTYPE
    assert KEYPOINT_DTYPE.itemsize == 28  # cv::KeyPoint


def test_header_cites_the_reference_interface():
    hdr = open(os.path.join(ROOT, "include", "sfm_match.h")).read()
    for cite in ("src/Sfm.cpp:590-608", "src/Sfm.cpp:511-515", "include/Sfm.h:60", "include/Utilities.h:27"):
        assert cite in hdr


def test_dmatch_layout_is_cv_dmatch():
    assert _lib.DMATCH_DTYPE.itemsize == 16
    assert [_lib.DMATCH_DTYPE.fields[k][1] for k in ("queryIdx", "trainIdx", "imgIdx", "distance")] == [0, 4, 8, 12]


def test_default_config_is_the_reference_constants():
    cfg = _lib.SfmmConfig()
    _lib.load().sfmm_default_config(C.byref(cfg))
    assert cfg.struct_size == C.sizeof(_lib.SfmmConfig)
    assert cfg.norm == _lib.NORM_L2 and cfg.cross_check == 0
    assert np.float32(cfg.ratio) == np.float32(0.8)


def test_row_pitch():
    L = _lib.load()
    assert L.sfmm_row_pitch(61, _lib.U8) == 64 and L.sfmm_row_pitch(32, _lib.U8) == 32
    assert L.sfmm_row_pitch(64, _lib.U8) == 64 and L.sfmm_row_pitch(65, _lib.U8) == 128
    assert L.sfmm_row_pitch(128, _lib.F32) == 512 and L.sfmm_row_pitch(129, _lib.U8) == 0


def test_bad_config_is_rejected_before_touching_cuda():
    L = _lib.load()
    cfg = _lib.SfmmConfig()
    L.sfmm_default_config(C.byref(cfg))
    cfg.norm = 7
    ctx = C.c_void_p()
    assert L.sfmm_create(C.byref(cfg), C.byref(ctx)) == _lib.SFMM_EINVAL
    assert b"norm" in L.sfmm_last_error(None)
    assert L.sfmm_create(None, C.byref(ctx)) == _lib.SFMM_EINVAL


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a CUDA device")
def test_no_gpu_means_error_not_fallback():
    with pytest.raises(SfmmError) as e:
        Matcher(_lib.NORM_HAMMING)
    assert e.value.code == _lib.SFMM_ENODEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sfm_danpipeline_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "liboracle" not in src and "bf_oracle" not in src, f


def test_every_kernel_header_is_tracked_by_the_build():
    """A .cuh missing from build.HEADERS lets a stale libsfmmatch.so travel to the GPU box unnoticed."""
    import glob
    import re
    from sfm_danpipeline_b200 import build
    csrc = build.CSRC
    included = set()
    for f in glob.glob(os.path.join(csrc, "*.cu")) + glob.glob(os.path.join(csrc, "*.cuh")):
        included |= set(re.findall(r'#include "([\w.]+\.cuh)"', open(f).read()))
    assert included and included <= set(build.HEADERS), included - set(build.HEADERS)
    assert not build.stale(), "libsfmmatch.so is older than its sources: run python -m sfm_danpipeline_b200.build"


def test_default_engines_are_auto():
    """AUTO = 0 for both selectors, so a zero-initialised SfmmConfig means "let the library pick"."""
    from sfm_danpipeline_b200 import _lib
    assert _lib.BINARY_AUTO == 0 and _lib.FLOAT_AUTO == 0
    cfg = _lib.SfmmConfig()
    _lib.load().sfmm_default_config(C.byref(cfg))
    assert cfg.binary_engine == _lib.BINARY_AUTO and cfg.float_mode == _lib.FLOAT_AUTO


@pytest.mark.skipif(_has_gpu(), reason="only meaningful without a CUDA device")
def test_group_without_gpu_is_an_error_too():
    from sfm_danpipeline_b200 import GroupMatcher
    with pytest.raises(SfmmError) as e:
        GroupMatcher(2, _lib.NORM_HAMMING)
    assert e.value.code == _lib.SFMM_ENODEVICE


def test_nccl_is_not_a_link_time_dependency():
    """The group API dlopens NCCL on first use: a single-GPU user never needs it installed."""
    import subprocess
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "nccl" not in out.lower()


def test_stats_mirror_matches_the_header():
    """The ctypes mirror of SfmmStats lists exactly the header's fields, in order (all 8-byte fields: no padding questions)."""
    import re
    text = open(os.path.join(ROOT, "include", "sfm_match.h")).read()
    body = re.search(r"typedef struct SfmmStats \{(.*?)\} SfmmStats;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(int64_t|double)\s+(\w+)\s*;", body)
    assert [n for _t, n in fields] == [n for n, _c in _lib.SfmmStats._fields_]
    for (t, n), (_n, c) in zip(fields, _lib.SfmmStats._fields_):
        assert c is (C.c_int64 if t == "int64_t" else C.c_double), n
    assert C.sizeof(_lib.SfmmStats) == 8 * len(fields)
    assert "tensor_kind" in dict(_lib.SfmmStats._fields_)


def test_f4x_digit_expansion_is_exact():
    """TM_F4X (csrc/float_tensor.cuh, binary_unpack4x_kernel) writes E = 512 - popc(t) into 17 spare elements as E2M1 digits against the
    query-side constants 6 x16, 1.  Restated here: every E in 0..512 is reproduced exactly from representable digits (the kernel itself is
    checked for every popcount by tests/test_parity_binary.py::test_f4x_key_term_covers_every_popcount on the GPU)."""
    e2m1 = {0: 0.0, 1: 0.5, 2: 1.0, 3: 1.5, 4: 2.0, 5: 3.0, 6: 4.0, 7: 6.0}  # nibble -> value
    def code(units):  # units of 0.5 -> nibble (the kernel's lambda)
        return units if units <= 4 else (5 if units == 6 else 6)
    for pc in range(0, 513):
        E = 512 - pc
        a, r = divmod(E, 36)
        b, c = divmod(r, 3)
        u = b if b <= 4 else ((4 if b == 5 else 6) if b <= 7 else 8)
        v = b - u
        assert a <= 14 and u in (0, 1, 2, 3, 4, 6, 8) and v in (0, 1, 2, 3) and c in (0, 1, 2)
        train = [7 if s < a else 0 for s in range(14)] + [code(u), code(v), 2 * c]
        query = [7] * 16 + [2]
        assert sum(e2m1[q] * e2m1[t] for q, t in zip(query, train)) == E, pc
