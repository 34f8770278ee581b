"""Where the end-to-end time goes (host pack + H2D / match / table export)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sfm_danpipeline_b200 import Matcher, synth
kind = sys.argv[1] if len(sys.argv) > 1 else "binary"
n_img = int(sys.argv[2]) if len(sys.argv) > 2 else (50 if kind == "binary" else 60)
n_desc = int(sys.argv[3]) if len(sys.argv) > 3 else (5000 if kind == "binary" else 8000)
if kind == "binary":
    descs, norm = synth.binary_images(n_img, n_desc, seed=0), 0
else:
    descs, norm = synth.float_images(n_img, n_desc, seed=0), 1
m = Matcher(norm)
for it in range(4):
    t0 = time.perf_counter(); m.set_descriptors(descs)
    t1 = time.perf_counter(); m.match_all_pairs()
    t2 = time.perf_counter(); tab = m.result_table()
    t3 = time.perf_counter()
    print(f"{kind} iter {it}: set_descriptors {1e3*(t1-t0):.2f} ms  match_all_pairs {1e3*(t2-t1):.2f} ms  result_table {1e3*(t3-t2):.2f} ms  "
          f"(device {m.stats()['last_match_ms']:.2f} ms, knn {m.stats()['last_knn_ms']:.2f} ms, matches {len(tab[3])})")
