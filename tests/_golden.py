"""Loader for the committed cv2 golden vectors (tests/golden/*.npz, made by make_golden.py)."""
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

SETS = {  # name -> (norm, descriptor dtype)
    "temple_akaze": (0, np.uint8),
    "temple_orb": (0, np.uint8),
    "synth_binary": (0, np.uint8),
    "temple_sift": (1, np.float32),
    "synth_float": (1, np.float32),
    # the reference's literal call for detectors 2 and 3: NORM_L2 over CV_8U rows (src/Sfm.cpp:593)
    "temple_orb_l2": (1, np.uint8),
    "temple_akaze_l2": (1, np.uint8),
}


class GoldenSet:
    """descs: list of per-image arrays; pairs: [(q, t, knn_dist, knn_idx, match_q, match_cross)]."""

    def __init__(self, name):
        self.name = name
        self.norm, dt = SETS[name]
        z = np.load(os.path.join(HERE, name + ".npz"))
        rows = z["rows"]
        offs = np.concatenate([[0], np.cumsum(rows)])
        desc = z["desc"].astype(dt)
        self.descs = [np.ascontiguousarray(desc[offs[i]:offs[i + 1]]) for i in range(len(rows))]
        self.pairs = []
        ko = mo = 0
        p = 0
        for q in range(len(rows) - 1):
            for t in range(q + 1, len(rows)):
                nq, nm = int(rows[q]), int(z["match_count"][p])
                self.pairs.append((q, t, z["knn_dist"][ko:ko + nq], z["knn_idx"][ko:ko + nq],
                                   z["match_q"][mo:mo + nm], z["match_cross"][mo:mo + nm]))
                ko += nq; mo += nm; p += 1

    def expected(self, pair_index, cross_check=False):
        """(queryIdx, trainIdx, distance) arrays cv2 produced for this pair."""
        q, t, kd, ki, mq, mx = self.pairs[pair_index]
        sel = mq[mx] if cross_check else mq
        return sel, ki[sel, 0], kd[sel, 0].astype(np.float32)
