"""Dev probe: one 640x480 frame through orbx_extract, repeated (latency of the drop-in's per-image call)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from multi_orb_slam_b200.extractor import ORBextractor
from multi_orb_slam_b200.synth import textured
img = textured(640, 480, 0)
ex = ORBextractor(1000, 1.2, 8, 20, 7, image_size=(640, 480), max_batch=1)
for _ in range(20):
    k, d = ex(img)
ts = []
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 200):
    t0 = time.perf_counter(); k, d = ex(img); ts.append(time.perf_counter() - t0)
print(f"single frame: median {np.median(ts) * 1e3:.3f} ms, min {min(ts) * 1e3:.3f} ms, {len(k)} keypoints")
