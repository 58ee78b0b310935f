"""Pins the oracle against the committed golden fixtures (tools/make_golden.py):
  * cvprim primitives == cv2 4.13.0 outputs (resize chain, REFLECT_101 border, 7x7 blur, FAST-9/16
    at both thresholds, scalar fastAtan2);
  * the restated extractor == the reference's own ORBextractor.cc compiled verbatim (keypoints,
    order, descriptors), on three geometries.
Runs without cv2 and without /root/reference."""
import ctypes as C
import hashlib
import os

import numpy as np
import pytest

from multi_orb_slam_b200.synth import textured

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def prim():
    return np.load(os.path.join(GOLD, "cv2_primitives.npz"))


@pytest.fixture(scope="module")
def lib(oracle_port):
    return oracle_port.load("port")


def test_resize_chain(lib, prim):
    cur = prim["image"]
    for l in range(1, 4):
        want = prim[f"resize_{l}"]
        got = np.zeros_like(want)
        lib.cvp_resize(C.c_void_p(cur.ctypes.data), C.c_size_t(cur.strides[0]), cur.shape[1], cur.shape[0],
                       C.c_void_p(got.ctypes.data), C.c_size_t(got.strides[0]), want.shape[1], want.shape[0])
        assert np.array_equal(got, want), f"level {l}"
        cur = np.ascontiguousarray(want)


def test_border(lib, prim):
    img, want = prim["image"], prim["border"]
    got = np.zeros_like(want)
    got[19:-19, 19:-19] = img
    lib.cvp_border(C.c_void_p(got.ctypes.data), C.c_size_t(got.strides[0]), img.shape[1], img.shape[0], 19)
    assert np.array_equal(got, want)


def test_blur(lib, prim):
    img = np.ascontiguousarray(prim["image"])
    got = np.zeros_like(img)
    lib.cvp_blur(C.c_void_p(img.ctypes.data), C.c_size_t(img.strides[0]), C.c_void_p(got.ctypes.data),
                 C.c_size_t(got.strides[0]), img.shape[1], img.shape[0])
    assert np.array_equal(got, prim["blur"])


@pytest.mark.parametrize("th", [7, 20])
def test_fast(lib, prim, th):
    img = np.ascontiguousarray(prim["image"])
    out = np.zeros((img.size, 3), dtype=np.int32)
    lib.cvp_fast.restype = C.c_int
    n = lib.cvp_fast(C.c_void_p(img.ctypes.data), C.c_size_t(img.strides[0]), img.shape[1], img.shape[0], th, 1,
                     C.c_void_p(out.ctypes.data), img.size)
    assert np.array_equal(out[:n], prim[f"fast_{th}"])


def test_fast_atan2(lib, prim):
    y, x = np.ascontiguousarray(prim["atan2_y"]), np.ascontiguousarray(prim["atan2_x"])
    got = np.zeros_like(x)
    lib.cvp_atan2(C.c_void_p(y.ctypes.data), C.c_void_p(x.ctypes.data), C.c_void_p(got.ctypes.data), len(x))
    assert np.array_equal(got, prim["atan2"])


@pytest.mark.parametrize("name", ["small", "vga", "kitti"])
def test_extractor_restatement_equals_verbatim_reference(oracle_port, name):
    g = np.load(os.path.join(GOLD, f"ref_extract_{name}.npz"))
    img = g["image"] if "image" in g.files else textured(int(g["width"]), int(g["height"]), int(g["seed"]))
    if hashlib.sha256(img.tobytes()).hexdigest() != str(g["image_sha256"]):
        pytest.skip("synthetic generator produced different bytes on this platform (numpy version?)")
    port = oracle_port.extractor("port", nfeatures=int(g["nfeatures"]))
    k, d, counts = port.extract(img)
    assert np.array_equal(counts, g["counts"])
    assert k.tobytes() == g["kps"].tobytes(), "keypoints (x, y, size, angle, response, octave) differ"
    assert np.array_equal(d, g["desc"])
    if "pyramid_l3" in g.files:
        assert np.array_equal(port.pyramid_level(3), g["pyramid_l3"])


def test_known_constants(oracle_port):
    """Hand-checkable values from the reference's constructor (src/ORBextractor.cc:411-471)."""
    port = oracle_port.extractor("port")
    assert list(port.features_per_level()) == [217, 181, 151, 126, 105, 87, 73, 60]
    assert list(oracle_port.extractor("port", nfeatures=500).features_per_level()) == [109, 90, 75, 63, 52, 44, 36, 31]
    assert list(oracle_port.extractor("port", nfeatures=2000).features_per_level()) == [434, 362, 302, 251, 209, 175, 145, 122]
    s = port.scale_tables()[0]
    assert s[0] == 1.0 and abs(s[1] - 1.2000000477) < 1e-9 and abs(s[7] - 3.5831816196) < 1e-6
    port.extract(textured(640, 480, 0))
    sizes = [(port.pyramid_level(l).shape[1] - 38, port.pyramid_level(l).shape[0] - 38) for l in range(8)]
    assert sizes == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]


def test_undistort_points_golden(oracle_port):
    """cv::undistortPoints as Frame::UndistortKeyPoints calls it (tools/make_golden_undistort.py)."""
    g = np.load(os.path.join(GOLD, "cv2_undistort.npz"))
    for i, c in enumerate(g["cams"]):
        got = oracle_port.undistort_points(g[f"pts_{i}"], c[0], c[1], c[2], c[3], c[4:9])
        assert np.array_equal(got.view(np.uint32), g[f"und_{i}"].view(np.uint32)), f"camera model {i}"
        # ComputeImageBounds from the four corners (src/Frame.cc:766-769)
        w, h = g[f"size_{i}"]
        u = g[f"und_{i}"][:4]
        want = (min(u[0, 0], u[2, 0]), max(u[1, 0], u[3, 0]), min(u[0, 1], u[1, 1]), max(u[2, 1], u[3, 1]))
        assert oracle_port.compute_image_bounds(int(w), int(h), c[0], c[1], c[2], c[3], c[4:9]) == tuple(float(v) for v in want)


def test_matcher_restatement_equals_reference_golden(oracle_port):
    """Outputs of the reference's own ORBmatcher.cc (verbatim build) recorded by tools/make_golden_matcher.py."""
    from golden_matcher_cases import CASES
    g = np.load(os.path.join(GOLD, "ref_matcher.npz"))
    for name, run in CASES.items():
        res = run(oracle_port, "port")
        for j, a in enumerate(res):
            assert np.array_equal(np.asarray(a), g[f"{name}_{j}"]), f"{name}[{j}]"
