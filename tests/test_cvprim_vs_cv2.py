"""Pins the oracle's OpenCV-primitive restatements (oracle/cvprim.cc) bit-exact against the real
cv2 4.13 wheel — the only OpenCV available here.  Skipped where cv2 is
not importable; tests/test_golden.py covers the same primitives from committed fixtures."""
import ctypes as C

import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from multi_orb_slam_b200.synth import textured


@pytest.fixture(scope="module")
def lib(oracle_port):
    return oracle_port.load("port")


def _images():
    rng = np.random.default_rng(123)
    yield "noise", rng.integers(0, 256, size=(480, 640), dtype=np.uint8)
    yield "textured", textured(640, 480, 3)
    yield "odd", rng.integers(0, 256, size=(67, 131), dtype=np.uint8)
    yield "kitti", textured(1241, 376, 5)


@pytest.mark.parametrize("name,img", list(_images()))
def test_resize_chain(lib, name, img):
    cur = img
    for _ in range(4):
        h, w = cur.shape
        dw, dh = int(np.rint(np.float32(w) / np.float32(1.2))), int(np.rint(np.float32(h) / np.float32(1.2)))
        want = cv2.resize(cur, (dw, dh), interpolation=cv2.INTER_LINEAR)
        got = np.zeros((dh, dw), dtype=np.uint8)
        lib.cvp_resize(C.c_void_p(cur.ctypes.data), C.c_size_t(cur.strides[0]), w, h,
                       C.c_void_p(got.ctypes.data), C.c_size_t(got.strides[0]), dw, dh)
        assert np.array_equal(got, want), name
        cur = want


@pytest.mark.parametrize("name,img", list(_images()))
def test_border(lib, name, img):
    want = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    got = np.zeros_like(want)
    got[19:-19, 19:-19] = img
    lib.cvp_border(C.c_void_p(got.ctypes.data), C.c_size_t(got.strides[0]), img.shape[1], img.shape[0], 19)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("name,img", list(_images()))
def test_gaussian_blur(lib, name, img):
    want = cv2.GaussianBlur(img, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    got = np.zeros_like(img)
    lib.cvp_blur(C.c_void_p(img.ctypes.data), C.c_size_t(img.strides[0]), C.c_void_p(got.ctypes.data),
                 C.c_size_t(got.strides[0]), img.shape[1], img.shape[0])
    assert np.array_equal(got, want)


def _fast(lib, img, th):
    cap = img.size
    out = np.zeros((cap, 3), dtype=np.int32)
    lib.cvp_fast.restype = C.c_int
    n = lib.cvp_fast(C.c_void_p(img.ctypes.data), C.c_size_t(img.strides[0]), img.shape[1], img.shape[0], th, 1,
                     C.c_void_p(out.ctypes.data), cap)
    return out[:n]


@pytest.mark.parametrize("th", [7, 20])
@pytest.mark.parametrize("name,img", list(_images()))
def test_fast(lib, name, img, th):
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True,
                                         type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    kps = det.detect(img, None)
    want = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], dtype=np.int32).reshape(-1, 3)
    got = _fast(lib, img, th)
    assert np.array_equal(got, want), (name, th, len(got), len(want))


def test_fast_cells(lib):
    """Per-cell sub-image calls, as the reference makes them (src/ORBextractor.cc:790-830)."""
    img = textured(640, 480, 9)
    det = cv2.FastFeatureDetector_create(threshold=20, nonmaxSuppression=True)
    for (y0, y1, x0, x1) in [(16, 54, 16, 53), (200, 238, 400, 437), (432, 464, 605, 624), (16, 23, 16, 23)]:
        sub = img[y0:y1, x0:x1]
        kps = det.detect(np.ascontiguousarray(sub), None)
        want = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], dtype=np.int32).reshape(-1, 3)
        got = _fast(lib, sub, 20)  # strided view: exercises step != width
        assert np.array_equal(got, want)


def test_fast_atan2(lib):
    rng = np.random.default_rng(5)
    y = rng.integers(-250000, 250000, size=100000).astype(np.float32)
    x = rng.integers(-250000, 250000, size=100000).astype(np.float32)
    y[:10] = 0
    x[5:15] = 0
    # the scalar cv::fastAtan2 the reference calls (ORBextractor.cc:103); cv2.phase's SIMD path
    # contracts to FMA and differs by 1 ulp on ~2 % of inputs, so it is NOT the pin
    want = np.array([cv2.fastAtan2(float(b), float(a)) for a, b in zip(x, y)], dtype=np.float32)
    got = np.zeros_like(x)
    lib.cvp_atan2(C.c_void_p(y.ctypes.data), C.c_void_p(x.ctypes.data), C.c_void_p(got.ctypes.data), len(x))
    assert np.array_equal(got, want)


def test_undistort_points(oracle_port):
    """cv::undistortPoints(src, dst, K, dist, Mat(), K) (Frame::UndistortKeyPoints, src/Frame.cc:692)."""
    rng = np.random.default_rng(5)
    for trial in range(40):
        fx, fy, cx, cy = (np.float32(v) for v in (rng.uniform(300, 900), rng.uniform(300, 900), rng.uniform(250, 700),
                                                  rng.uniform(180, 400)))
        dist = np.array([rng.uniform(-0.4, 0.4), rng.uniform(-0.3, 0.3), rng.uniform(-0.01, 0.01), rng.uniform(-0.01, 0.01),
                         rng.uniform(-0.2, 0.2) if trial % 2 else 0], dtype=np.float32)
        pts = np.stack([rng.uniform(-50, 1330, 300), rng.uniform(-50, 770, 300)], axis=1).astype(np.float32)
        K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1]], dtype=np.float32)
        want = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, dist if trial % 2 else dist[:4], None, K).reshape(-1, 2)
        got = oracle_port.undistort_points(pts, fx, fy, cx, cy, dist)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"trial {trial}"


def test_small_matrix_algebra_matches_cv_gemm(oracle_port):
    """cv::Mat expressions of the pose-based searches: `R*x + t` takes cv::gemm's float32 small-matrix path,
    `-R.t()*x` (camera centres, src/ORBmatcher.cc:595, 2235, 3755, 4285) the double-accumulating generic one."""
    lib = oracle_port.load("port")
    lib.om_gemm3_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_void_p]
    rng = np.random.default_rng(9)
    for trial in range(500):
        A = rng.normal(0, 1, (3, 3)).astype(np.float32)
        x = rng.normal(0, 5, (3, 1)).astype(np.float32)
        c = rng.normal(0, 5, (3, 1)).astype(np.float32)
        out = np.zeros(3, np.float32)
        lib.om_gemm3_probe(A.ctypes.data, x.ctypes.data, c.ctypes.data, 1.0, 0, out.ctypes.data)
        assert np.array_equal(out, cv2.gemm(A, x, 1, c, 1)[:, 0]), "R*x + t"
        lib.om_gemm3_probe(A.ctypes.data, x.ctypes.data, None, -1.0, 0, out.ctypes.data)
        assert np.array_equal(out, cv2.gemm(A, x, -1, None, 0)[:, 0]), "-R*x"
        lib.om_gemm3_probe(A.ctypes.data, x.ctypes.data, None, -1.0, 1, out.ctypes.data)
        assert np.array_equal(out, cv2.gemm(A, x, -1, None, 0, flags=cv2.GEMM_1_T)[:, 0]), "-R.t()*x"
