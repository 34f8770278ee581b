"""The C++ side of the drop-in: include/sfm_match_opencv.hpp compiled against a mock <opencv2/core.hpp>
(OpenCV's C++ headers are not in this image) and linked to libsfmmatch.so."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from sfm_danpipeline_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "adapter_test.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "adapter_test")


def _build():
    lib_dir = os.path.dirname(_lib.LIB_PATH)
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(SRC), os.path.getmtime(
            os.path.join(ROOT, "include", "sfm_match_opencv.hpp"))):
        subprocess.check_call(["g++", "-std=c++11", "-O1", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"),
                               "-I", os.path.join(ROOT, "tests", "cpp", "mock_opencv"), SRC, "-o", EXE,
                               "-L", lib_dir, "-lsfmmatch", "-Wl,-rpath," + lib_dir])
    return EXE


def test_adapter_compiles_as_cxx11_against_the_abi():
    _lib.load()
    _build()


def _write_case(path, descs, norm, cross):
    with open(path, "wb") as f:
        f.write(np.array([len(descs), descs[0].shape[1], int(descs[0].dtype == np.float32), int(cross), norm], np.int32).tobytes())
        f.write(np.array([d.shape[0] for d in descs], np.int32).tobytes())
        for d in descs:
            f.write(np.ascontiguousarray(d).tobytes())
        for q, t in synth.all_pairs(len(descs)):
            m = oracle.match_pair(descs[q], descs[t], norm, 0.8, cross)
            f.write(np.int32(len(m)).tobytes())
            f.write(m.tobytes())


@pytest.mark.gpu
@pytest.mark.parametrize("kind,cross", [("binary", False), ("binary", True), ("float", False), ("binary_l2", False)])
def test_patched_getmatching_equals_oracle(tmp_path, kind, cross):
    """binary_l2 = the default-constructed adapter (cv::NORM_L2, like src/Sfm.cpp:593) over CV_8U Mats."""
    exe = _build()
    if kind.startswith("binary"):
        descs, norm = synth.binary_images(4, [600, 0, 333, 1030], seed=21), (1 if kind == "binary_l2" else 0)
        descs[1] = np.zeros((0, 61), np.uint8)
    else:
        descs, norm = synth.float_images(3, [300, 200, 150], seed=22), 1
    case = str(tmp_path / "case.bin")
    _write_case(case, descs, norm, cross)
    import torch
    n_dev = torch.cuda.device_count()  # every visible GPU: the group broadcasts with NCCL when there is more than one
    z = np.load(os.path.join(ROOT, "tests", "golden", "temple_orb_features.npz"))
    img_file = str(tmp_path / "image.bin")
    with open(img_file, "wb") as f:  # sfmm::OrbExtractor on the reference's first temple image: cv::ORB finds 500 keypoints there
        f.write(np.array([480, 640, int(z["counts"][0])], np.int32).tobytes())
        f.write(np.ascontiguousarray(z["images"][0]).tobytes())
    r = subprocess.run([exe, case, str(n_dev), img_file], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "adapter ok" in r.stdout and f"multi-gpu ok on {n_dev} device(s)" in r.stdout and "orb ok: 500 keypoints" in r.stdout
