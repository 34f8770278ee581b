"""Small all-engines workload for compute-sanitizer (race detection / sanitizers).

    compute-sanitizer --tool memcheck  python tools/sanitizer_workload.py
    compute-sanitizer --tool racecheck python tools/sanitizer_workload.py

Covers the POPC kernel, the tensor kernels (fp4 packed / fp4 with the key term in the MMA / i8 packed / i8 / fp16 / TF32 exact /
TF32 rank+collect+refine, TMEM-A and shared-memory-A forms, plain and threshold-skipping epilogues), the exact fp32 kernel, cross-check on and off,
empty / tiny / ragged images, single-pair and raw-knn calls; every result is compared with the oracle."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle  # noqa: E402
from sfm_danpipeline_b200 import BINARY_POPC, BINARY_TENSOR, FLOAT_AUTO, FLOAT_EXACT, FLOAT_TENSOR, Matcher, synth  # noqa: E402

d = synth.binary_images(4, [300, 130, 0, 257], seed=3)
d[2] = np.zeros((0, 61), np.uint8)
for cross in (False, True):
    for eng in (BINARY_POPC, BINARY_TENSOR):
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
# 512-bit descriptors: the tensor engine's 32-bit-key path (TM_I8); 256-bit: packed keys with two K-blocks
rng = np.random.default_rng(5)
for cols in (64, 32):
    w = [rng.integers(0, 256, (n, cols), dtype=np.uint8) for n in (150, 129, 40)]
    with Matcher(0, 0.8, False, binary_engine=BINARY_TENSOR) as m:
        m.set_descriptors(w)
        m.match_all_pairs()
        for (q, t) in synth.all_pairs(3):
            assert m.getMatching(q, t).tobytes() == oracle.match_pair(w[q], w[t], 0, 0.8, False).tobytes()
# arbitrary floats: TF32 rank + collect + refine (float_path 3) equals the exact kernel bit for bit; 96-d: the TF32 exact path
x = synth.float_images(3, [200, 129, 70], seed=6, integer=False)
with Matcher(1, 0.8, False, float_mode=FLOAT_AUTO) as m, Matcher(1, 0.8, False, float_mode=FLOAT_EXACT) as mx:
    for mm in (m, mx):
        mm.set_descriptors(x)
        mm.match_all_pairs()
    assert m.stats()["float_path"] == 3
    for (q, t) in synth.all_pairs(3):
        assert m.getMatching(q, t).tobytes() == mx.getMatching(q, t).tobytes()
g = [np.floor(rng.random((n, 96), dtype=np.float32) * 100).astype(np.float32) for n in (140, 129)]
with Matcher(1, 0.8, False, float_mode=FLOAT_TENSOR) as m:
    m.set_descriptors(g)
    m.match_all_pairs()
    assert m.getMatching(0, 1).tobytes() == oracle.match_pair(g[0], g[1], 1, 0.8, False).tobytes()
# round 2: L2 over CV_8U rows (widened on the device by ingest_rows_kernel), cross-check through candidate columns on the float /
# arbitrary-float paths (exact kernel in reverse + gather mode), the key-term-in-the-MMA float kernel, ORB extraction
u = [rng.integers(0, 256, (n, 61), dtype=np.uint8) for n in (140, 129, 3)]
for cross in (False, True):
    with Matcher(1, 0.9, cross) as m:
        m.set_descriptors(u)
        m.match_all_pairs()
        for (q, t) in synth.all_pairs(3):
            assert m.getMatching(q, t).tobytes() == oracle.match_pair(u[q], u[t], 1, 0.9, cross).tobytes()
with Matcher(1, 0.8, True, float_mode=FLOAT_AUTO) as m, Matcher(1, 0.8, True, float_mode=FLOAT_EXACT) as mx:
    for mm in (m, mx):
        mm.set_descriptors(x)
        mm.match_all_pairs()
    assert m.stats()["float_path"] == 3
    for (q, t) in synth.all_pairs(3):
        assert m.getMatching(q, t).tobytes() == mx.getMatching(q, t).tobytes()
from oracle import orb_oracle  # noqa: E402
from sfm_danpipeline_b200 import OrbExtractor  # noqa: E402
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "temple_orb_features.npz"))
img = np.ascontiguousarray(z["images"][1][100:330, 150:470])
with OrbExtractor(0) as orb:
    kp, desc = orb.detectAndCompute(img)
okp, od = orb_oracle.detect_and_compute(img)
assert {(int(k["octave"]), float(k["x"]), float(k["y"])): bytes(dd) for k, dd in zip(kp, desc)} == \
       {(int(k["octave"]), float(k["x"]), float(k["y"])): bytes(dd) for k, dd in zip(okp, od)}
# second half of round 2: the FP4 pipe -- TM_F4P is what the binary sets above ran on; TM_F4X (key term in the MMA, threshold-skipping epilogue with
# shared-memory atomics between the groups; forced, the sets are small) with two and three epilogue groups, kind::i8, and the plain fold of TM_F16X
for env in ({"SFMM_F4X": "1"}, {"SFMM_F4X": "1", "SFMM_EPI_GROUPS": "2"}, {"SFMM_NO_F4": "1"}, {"SFMM_NO_SKIP": "1"}, {"SFMM_EPI_GROUPS": "3"}):
    os.environ.update(env)
    for cross in (False, True):
        with Matcher(0, 0.8, cross) as m:
            m.set_descriptors(d)
            m.match_all_pairs()
            for (q, t) in synth.all_pairs(4):
                assert m.getMatching(q, t).tobytes() == oracle.match_pair(d[q], d[t], 0, 0.8, cross).tobytes()
        with Matcher(1, 0.8, cross) as m:
            m.set_descriptors(f)
            m.match_all_pairs()
            for (q, t) in synth.all_pairs(3):
                assert m.getMatching(q, t).tobytes() == oracle.match_pair(f[q], f[t], 1, 0.8, cross).tobytes()
    for k in env:
        del os.environ[k]
print("sanitizer workload ok")
