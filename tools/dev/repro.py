import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from multi_orb_slam_b200.extractor import ORBextractor
w, h, nf, seed, mb = 1241, 376, 1733, 910432957, int(sys.argv[1]) if len(sys.argv) > 1 else 3
r = np.random.default_rng(seed)
img = (128 + r.integers(-5, 6, (h, w))).astype(np.uint8)
for _ in range(6):
    y, x = int(r.integers(0, h - 40)), int(r.integers(0, w - 40))
    img[y:y + 30, x:x + 30] += np.uint8(r.integers(20, 90))
ex = ORBextractor(nf, 1.2, 8, 20, 7, image_size=(w, h), max_batch=mb)
k, d = ex(img)
print("ok", len(k))
