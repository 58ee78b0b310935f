#!/usr/bin/env python
"""Golden vectors of cv2.undistortPoints (cv2 4.13.0) for tests/test_golden.py: the call
Frame::UndistortKeyPoints makes (src/Frame.cc:692), `cv::undistortPoints(mat, mat, mK, mDistCoef, Mat(), mK)`,
on seeded points for several camera models.  Writes tests/golden/cv2_undistort.npz."""
import os

import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
rng = np.random.default_rng(2024)
cams = [  # fx, fy, cx, cy, k1, k2, p1, p2, k3
    (517.306408, 516.469215, 318.643040, 255.313989, 0.262383, -0.953104, -0.005358, 0.002628, 1.163314),  # TUM1.yaml
    (520.908620, 521.007327, 325.141442, 249.701764, 0.231222, -0.784899, -0.003257, -0.000105, 0.917205),  # TUM2.yaml
    (458.654, 457.296, 367.215, 248.375, -0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0),      # EuRoC.yaml
    (700.0, 705.0, 640.0, 360.0, -0.35, 0.12, 0.001, -0.002, -0.02),
]
out = {"cams": np.array(cams, dtype=np.float32)}
for i, c in enumerate(cams):
    c = np.array(c, dtype=np.float32)
    K = np.array([[c[0], 0, c[2]], [0, c[1], c[3]], [0, 0, 1]], dtype=np.float32)
    w, h = (1280, 720) if i == 3 else (752, 480) if i == 2 else (640, 480)
    pts = np.stack([rng.uniform(0, w, 400), rng.uniform(0, h, 400)], axis=1).astype(np.float32)
    pts[:4] = [[0, 0], [w, 0], [0, h], [w, h]]  # the corners ComputeImageBounds undistorts (src/Frame.cc:749-757)
    dist = c[4:9] if c[8] != 0 else c[4:8]      # four-parameter model when k3 is absent, like the yaml files
    und = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, dist, None, K).reshape(-1, 2)
    out[f"pts_{i}"], out[f"und_{i}"] = pts, und.astype(np.float32)
    out[f"size_{i}"] = np.array([w, h], dtype=np.int32)
np.savez_compressed(os.path.join(OUT, "cv2_undistort.npz"), **out)
print("wrote cv2_undistort.npz with cv2", cv2.__version__)
