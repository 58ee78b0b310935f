#!/usr/bin/env python
"""Generates tests/golden/*.npz (authoring container only).

Sources of truth:
  * cv2 4.13.0 (the only OpenCV available; the reference pins 'OpenCV >= 2.4.3') for the pixel
    primitives the reference calls: resize, copyMakeBorder, GaussianBlur, FAST, fastAtan2;
  * the reference's own src/ORBextractor.cc compiled verbatim (oracle/_ref/liborb_ref.so, built
    by oracle/Makefile from /root/reference) for end-to-end keypoints + descriptors.
The fixtures travel to the GPU box, where neither cv2 nor /root/reference may be assumed.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cv2  # noqa: E402
import oracle_lib as O  # noqa: E402
from multi_orb_slam_b200.synth import textured  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
O.build_oracle()
assert O.load("ref") is not None, "oracle/_ref/liborb_ref.so missing: /root/reference not mounted?"

# ---- cv2 primitives on a small image ---------------------------------------------------------
img = textured(200, 150, 21)
prim = {"image": img}
cur = img
for l in range(1, 4):
    h, w = cur.shape
    dw, dh = int(np.rint(np.float32(w) / np.float32(1.2))), int(np.rint(np.float32(h) / np.float32(1.2)))
    cur = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
    prim[f"resize_{l}"] = cur
prim["border"] = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
prim["blur"] = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
for th in (7, 20):
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = det.detect(img, None)
    prim[f"fast_{th}"] = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], dtype=np.int32).reshape(-1, 3)
rng = np.random.default_rng(9)
ay = rng.integers(-200000, 200000, size=4000).astype(np.float32)
ax = rng.integers(-200000, 200000, size=4000).astype(np.float32)
ay[:8] = 0
ax[4:12] = 0
prim["atan2_y"], prim["atan2_x"] = ay, ax
prim["atan2"] = np.array([cv2.fastAtan2(float(b), float(a)) for a, b in zip(ax, ay)], dtype=np.float32)
np.savez_compressed(os.path.join(OUT, "cv2_primitives.npz"), **prim)

# ---- verbatim reference extractor ------------------------------------------------------------
cases = {"small": dict(size=(320, 240), seed=31, nfeatures=300), "vga": dict(size=(640, 480), seed=0, nfeatures=1000),
         "kitti": dict(size=(1241, 376), seed=5, nfeatures=2000)}
for name, c in cases.items():
    im = textured(c["size"][0], c["size"][1], c["seed"])
    ref = O.extractor("ref", nfeatures=c["nfeatures"])
    k, d, counts = ref.extract(im)
    out = dict(kps=k, desc=d, counts=counts, nfeatures=c["nfeatures"], seed=c["seed"], width=c["size"][0],
               height=c["size"][1], image_sha256=hashlib.sha256(im.tobytes()).hexdigest())
    if name == "small":
        out["image"] = im  # full bytes only for the small case; the others are regenerated and checksummed
        out["pyramid_l3"] = ref.pyramid_level(3)
    np.savez_compressed(os.path.join(OUT, f"ref_extract_{name}.npz"), **out)
    print(name, len(k), counts)
print("fixtures written to", OUT, [(f, os.path.getsize(os.path.join(OUT, f))) for f in sorted(os.listdir(OUT))])
