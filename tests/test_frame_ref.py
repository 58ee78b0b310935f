"""Pins the Frame-glue restatements (and through them the GPU kernels of multi_orb_slam_b200/frame.py) to the
reference's OWN src/Frame.cc, compiled verbatim with its real include/Frame.h into oracle/_ref/libframe_ref.so:
frames are built by the reference's two-camera RGB-D constructor (src/Frame.cc:148-346), so UndistortKeyPoints,
ComputeImageBounds, ComputeStereoFromRGBD, the camera index maps, AssignFeaturesToGrid and GetFeaturesInArea run
as written; SearchForInitialization of the verbatim ORBmatcher.cc then searches through that Frame.
Skipped where the verbatim build is absent."""
import numpy as np
import pytest

import oracle_lib as O
from multi_orb_slam_b200.synth import KP_DTYPE

pytestmark = pytest.mark.skipif(O.load("fref") is None, reason="oracle/_ref/libframe_ref.so not built (no /root/reference)")
TUM1 = (517.306408, 516.469215, 318.643040, 255.313989)
TUM1_DIST = (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)


def _keys(n, seed, w=640, h=480):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, KP_DTYPE)
    k["x"], k["y"] = rng.uniform(0, w - 1, n), rng.uniform(0, h - 1, n)
    k["octave"], k["angle"], k["size"], k["response"] = rng.integers(0, 8, n), rng.uniform(0, 360, n), 31, rng.uniform(1, 99, n)
    return k


@pytest.mark.parametrize("dist", [TUM1_DIST, TUM1_DIST[:4], (-0.28, 0.07, 0.0002, 0.00002, 0.0)])
def test_frame_constructor_glue(dist):
    rng = np.random.default_rng(5)
    k0, k1 = _keys(1000, 1), _keys(500, 2)
    d0, d1 = rng.integers(0, 256, (1000, 32), dtype=np.uint8), rng.integers(0, 256, (500, 32), dtype=np.uint8)
    depth = [np.where(rng.random((480, 640)) < 0.7, rng.uniform(0.4, 8, (480, 640)), 0).astype(np.float32) for _ in range(2)]
    ref = O.frame_glue_ref(k0, d0, k1, d1, depth[0], depth[1], 640, 480, *TUM1, dist, 40.0)
    d5 = list(dist) + [0.0] * (5 - len(dist))
    # restatement, per camera then concatenated like the reference's *_total arrays
    un = [O.undistort_keypoints(k, *TUM1, d5) for k in (k0, k1)]
    k_un = np.concatenate(un)
    assert ref["k_un"].tobytes() == k_un.tobytes()
    bounds = O.compute_image_bounds(640, 480, *TUM1, d5)
    assert ref["bounds"] == bounds
    st = [O.compute_stereo_from_rgbd(k, u, z, 40.0) for k, u, z in zip((k0, k1), un, depth)]
    assert np.array_equal(ref["uright"], np.concatenate([s[0] for s in st]))
    assert np.array_equal(ref["depth"], np.concatenate([s[1] for s in st]))
    cam = np.concatenate([np.zeros(1000, np.int32), np.ones(500, np.int32)])
    assert np.array_equal(ref["cam"], cam) and np.array_equal(ref["local"], np.concatenate([np.arange(1000), np.arange(500)]))
    # AssignFeaturesToGrid: per-camera grids of global indices
    for c in range(2):
        sel = np.nonzero(cam == c)[0]
        start, items = O.assign_features_to_grid(k_un[sel], bounds)
        assert np.array_equal(ref["grid_start"][c], start), f"camera {c}: cell starts"
        assert np.array_equal(ref["grid_items"][c, : start[-1]], sel[items]), f"camera {c}: cell contents"


def test_get_features_in_area():
    k0, k1 = _keys(1200, 3), _keys(600, 4)
    k0["x"][:4], k0["y"][:4] = [0, 639.6, 320, 5], [0, 479.6, 479.9, 240]
    rng = np.random.default_rng(6)
    cam = np.concatenate([np.zeros(1200, np.int32), np.ones(600, np.int32)])
    kk = np.concatenate([O.undistort_keypoints(k, *TUM1, TUM1_DIST) for k in (k0, k1)])
    bounds = O.compute_image_bounds(640, 480, *TUM1, TUM1_DIST)
    n_hit = 0
    for _ in range(60):
        c = int(rng.integers(0, 2))
        x, y = float(rng.uniform(bounds[0] - 30, bounds[1] + 30)), float(rng.uniform(bounds[2] - 30, bounds[3] + 30))
        r = float(rng.choice([3.0, 15.0, 40.0, 100.0, 900.0]))
        lo, hi = [(-1, -1), (0, 0), (2, 5), (3, -1), (0, 7)][int(rng.integers(0, 5))]
        ref = O.features_in_area_ref(k0, k1, 640, 480, *TUM1, TUM1_DIST, c, x, y, r, lo, hi)
        sel = np.nonzero(cam == c)[0]
        got = O.features_in_area(kk["x"][sel], kk["y"][sel], kk["octave"][sel], bounds, x, y, r, lo, hi)
        assert np.array_equal(ref, sel[got]), (c, x, y, r, lo, hi)
        n_hit += len(ref)
    assert n_hit > 1000


@pytest.mark.parametrize("seed,window,check_ori", [(0, 100, True), (1, 30, True), (2, 100, False)])
def test_search_for_initialization_through_reference_frames(seed, window, check_ori):
    from test_gpu_matcher import _frame_pair
    k1, d1, k2, d2 = _frame_pair(O, seed)
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
    b = O.search_for_initialization_frame_ref(k1, d1, k2, d2, 640, 480, *TUM1, TUM1_DIST, prev, window, 0.9, check_ori)
    # the restatement sees what the reference's matcher reads from its Frame: mvKeysUn and the undistorted bounds
    u1, u2 = O.undistort_keypoints(k1, *TUM1, TUM1_DIST), O.undistort_keypoints(k2, *TUM1, TUM1_DIST)
    bounds = O.compute_image_bounds(640, 480, *TUM1, TUM1_DIST)
    a = O.search_for_initialization(u1, d1, u2, d2, bounds, prev, window, 0.9, check_ori)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[0] > 20


@pytest.mark.parametrize("seed,size,nfeatures,disp,mbf,mb", [
    (3, (640, 480), 1000, (4, 9, 17, 30), 40.0, 0.08),
    (5, (400, 300), 300, (6, 13), 35.0, 0.09),
    (900, (1241, 376), 2000, (5, 12, 26, 44), 386.1448, 386.1448 / 718.856),  # KITTI-shaped (configs[3])
])
def test_compute_stereo_matches_vs_the_reference_text(seed, size, nfeatures, disp, mbf, mb):
    """Frame::ComputeStereoMatches is commented out in the fork (src/Frame.cc:782-956): oracle/gen_stereo_ref.py
    un-comments the reference's own text at build time into oracle/_ref/ and it runs here, as a member of the reference's
    real Frame class, against the restatement that checks the CUDA kernel — mvuRight / mvDepth bit for bit."""
    from test_frame_glue_oracle import _stereo_oracle
    _, _, ((kl, dl, pl, tl), (kr, dr, pr, _)) = _stereo_oracle(O, seed, nfeatures=nfeatures, W=size[0], H=size[1], disparities=disp)
    a = O.compute_stereo_matches(kl, dl, kr, dr, pl, pr, tl[0], tl[1], mbf, mb)
    b = O.compute_stereo_matches(kl, dl, kr, dr, pl, pr, tl[0], tl[1], mbf, mb, impl="ref")
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert (a[0] >= 0).sum() > 0.3 * len(kl)
