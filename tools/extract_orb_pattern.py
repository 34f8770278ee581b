#!/usr/bin/env python
"""Reads ORB's learned sampling pattern (Rublee et al. 2011; `bit_pattern_31_` in OpenCV's features2d/src/orb.cpp, 256 x 4 int32:
x0, y0, x1, y1 of every binary test) out of the cv2 wheel's binary and stores it as sfm_danpipeline_b200/orb_pattern.npy.
The table is a constant of the published method; it is found by its first twelve entries.  Run in the build container only."""
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    import cv2
    so = glob.glob(os.path.join(os.path.dirname(cv2.__file__), "*.so"))[0]
    data = open(so, "rb").read()
    head = np.array([8, -3, 9, 5, 4, 2, 7, -12, -11, 9, -8, 2], np.int32).tobytes()
    i = data.find(head)
    assert i >= 0 and data.find(head, i + 1) < 0, "pattern not found exactly once"
    tab = np.frombuffer(data[i:i + 256 * 4 * 4], np.int32).reshape(256, 4).copy()
    assert np.abs(tab).max() <= 15 and (tab[-1] == [-1, -6, 0, -11]).all()
    out = os.path.join(ROOT, "sfm_danpipeline_b200", "orb_pattern.npy")
    np.save(out, tab)
    print(out, tab.shape, "from cv2", cv2.__version__)


if __name__ == "__main__":
    sys.exit(main())
