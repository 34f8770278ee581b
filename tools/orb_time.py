import os, sys, time, numpy as np
sys.path.insert(0, os.getcwd())
from sfm_danpipeline_b200 import OrbExtractor
z = np.load("tests/golden/temple_orb_features.npz")
imgs = [np.ascontiguousarray(z["images"][i]) for i in range(len(z["images"]))]
with OrbExtractor(0) as orb:
    for im in imgs: orb.detectAndCompute(im)
    t0 = time.perf_counter(); n = 0
    for rep in range(20):
        for im in imgs:
            orb.detectAndCompute(im); n += 1
    dt = time.perf_counter() - t0
    print(os.environ.get("SFMM_ORB_NO_GRAPH", "graph"), "images/s %.0f" % (n / dt), "ms/image e2e %.3f" % (dt / n * 1e3), "device ms %.3f" % orb.stats()["last_ms"] if hasattr(orb, "stats") else "")
