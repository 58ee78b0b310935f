"""CPU checks of the oracle's Frame-glue restatements (src/Frame.cc:348-395, 673-706, 743-779, 959-985)
against direct numpy statements of the same lines."""
import math

import numpy as np

from multi_orb_slam_b200.synth import KP_DTYPE


def _keys(n, seed, w=640, h=480):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, KP_DTYPE)
    k["x"], k["y"] = rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n)
    k["octave"], k["angle"], k["size"], k["response"] = rng.integers(0, 8, n), rng.uniform(0, 360, n), 31, rng.uniform(1, 99, n)
    return k


def test_undistort_keypoints_carries_fields_and_identity(oracle_port):
    k = _keys(200, 1)
    cam = (517.3, 516.5, 318.6, 255.3)
    un = oracle_port.undistort_keypoints(k, *cam, [0.26, -0.95, -0.005, 0.0026, 1.16])
    for f in ("size", "angle", "response", "octave"):
        assert np.array_equal(un[f], k[f])
    want = oracle_port.undistort_points(np.stack([k["x"], k["y"]], axis=1), *cam, [0.26, -0.95, -0.005, 0.0026, 1.16])
    assert np.array_equal(un["x"], want[:, 0]) and np.array_equal(un["y"], want[:, 1])
    assert not np.array_equal(un["x"], k["x"])
    same = oracle_port.undistort_keypoints(k, *cam, [0, 0.3, 0.1, 0.1, 0])  # k1 == 0: mvKeysUn = mvKeys (:675-679)
    assert same.tobytes() == k.tobytes()


def test_compute_image_bounds_without_distortion(oracle_port):
    assert oracle_port.compute_image_bounds(640, 480, 500, 500, 320, 240, [0, 0, 0, 0, 0]) == (0.0, 640.0, 0.0, 480.0)


def test_compute_stereo_from_rgbd(oracle_port):
    k = _keys(300, 2)
    un = oracle_port.undistort_keypoints(k, 517.3, 516.5, 318.6, 255.3, [0.26, -0.95, -0.005, 0.0026, 1.16])
    rng = np.random.default_rng(3)
    depth = np.where(rng.random((480, 640)) < 0.7, rng.uniform(0.4, 8, (480, 640)), 0).astype(np.float32)
    ur, dz = oracle_port.compute_stereo_from_rgbd(k, un, depth, 40.0)
    for i in range(len(k)):
        d = depth[int(k["y"][i]), int(k["x"][i])]
        if d > 0:
            assert dz[i] == d and ur[i] == np.float32(un["x"][i] - np.float32(np.float32(40.0) / d))
        else:
            assert dz[i] == -1 and ur[i] == -1
    assert (dz > 0).sum() > 100 and (dz < 0).sum() > 20


def test_assign_features_to_grid(oracle_port):
    k = _keys(1500, 4)
    k["x"][:5], k["y"][:5] = [-3, 700, 10, 639.9, 0], [10, 10, -2, 479.9, 0]  # outside / on the edges
    bounds = (-8.5, 651.25, -6.0, 489.5)
    start, items = oracle_port.assign_features_to_grid(k, bounds)
    inv_w = np.float32(64) / (np.float32(bounds[1]) - np.float32(bounds[0]))
    inv_h = np.float32(48) / (np.float32(bounds[3]) - np.float32(bounds[2]))
    cells = {}
    for i in range(len(k)):
        px = int(math.floor(np.float32((k["x"][i] - np.float32(bounds[0])) * inv_w) + 0.5))
        py = int(math.floor(np.float32((k["y"][i] - np.float32(bounds[2])) * inv_h) + 0.5))
        if 0 <= px < 64 and 0 <= py < 48:
            cells.setdefault(px * 48 + py, []).append(i)
    assert start[-1] == sum(len(v) for v in cells.values()) == len(items)
    for c in range(64 * 48):
        assert list(items[start[c]:start[c + 1]]) == cells.get(c, [])
