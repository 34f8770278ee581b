"""Small all-engines workload for compute-sanitizer (race detection / sanitizers).

    compute-sanitizer --tool memcheck  python tools/sanitizer_workload.py
    compute-sanitizer --tool racecheck python tools/sanitizer_workload.py

Covers the POPC kernel, the i8 and TF32 tensor kernels, the exact fp32 kernel, cross-check on and off,
empty / tiny / ragged images, single-pair and raw-knn calls; every result is compared with the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from sfm_danpipeline_b200 import BINARY_TENSOR, FLOAT_EXACT, FLOAT_TENSOR, Matcher, synth  # noqa: E402

d = synth.binary_images(4, [300, 130, 0, 257], seed=3)
d[2] = np.zeros((0, 61), np.uint8)
for cross in (False, True):
    for eng in (0, BINARY_TENSOR):
        with Matcher(0, 0.8, cross, binary_engine=eng) as m:
            m.set_descriptors(d)
            m.match_all_pairs()
            for (q, t) in synth.all_pairs(4):
                assert m.getMatching(q, t).tobytes() == oracle.match_pair(d[q], d[t], 0, 0.8, cross).tobytes()
            m.knn_pair(0, 1)
            m.match_pair(3, 0)
f = synth.float_images(3, [200, 129, 70], seed=4)
for mode in (FLOAT_TENSOR, FLOAT_EXACT):
    for cross in (False, True):
        with Matcher(1, 0.8, cross, float_mode=mode) as m:
            m.set_descriptors(f)
            m.match_all_pairs()
            for (q, t) in synth.all_pairs(3):
                assert m.getMatching(q, t).tobytes() == oracle.match_pair(f[q], f[t], 1, 0.8, cross).tobytes()
print("sanitizer workload ok")
