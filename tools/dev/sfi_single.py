"""Dev probe: latency of one SearchForInitialization call through the host entry (the per-frame call of
Tracking::MonocularInitialization) on two consecutive synthetic frames."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from multi_orb_slam_b200.extractor import ORBextractor
from multi_orb_slam_b200.matcher import Frame, ORBmatcher
from multi_orb_slam_b200.synth import shifted_noisy, textured
img1 = textured(640, 480, 0); img2 = shifted_noisy(img1, 1000)
ex = ORBextractor(1000, 1.2, 8, 20, 7, image_size=(640, 480), max_batch=1)
(k1, d1), (k2, d2) = ex(img1), ex(img2)
m = ORBmatcher(0.9, True)
f1, f2 = Frame(k1, d1, 640, 480), Frame(k2, d2, 640, 480)
prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
for _ in range(20):
    n, m12 = m.SearchForInitialization(f1, f2, prev.copy(), 100)
ts = []
for _ in range(200):
    p = prev.copy(); t0 = time.perf_counter(); n, m12 = m.SearchForInitialization(f1, f2, p, 100); ts.append(time.perf_counter() - t0)
print(f"SearchForInitialization single pair: median {np.median(ts) * 1e3:.3f} ms, {n} matches")
