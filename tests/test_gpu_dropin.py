"""The reference's OWN classes with the GPU members: ORB_SLAM2::ORBmatcher / ORBextractor as declared by the reference's
headers, defined by multi_orb_slam_b200/dropin/*.cc over the C ABI, driven through the same C harness that drives the
verbatim reference build (oracle/matcher_ref_capi.cc; object graphs of Frame / KeyFrame / MapPoint built from flat
arrays) and compared with the oracle on the scenes of tests/test_matcher_ref.py.  The libraries are built where
/root/reference exists (tests/native/Makefile) and travel to the GPU box."""
import numpy as np
import pytest

import oracle_lib as O
from test_gpu_matcher import CALIB, CAM, _frame_pair, _projection_case, _rig_scene

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(O.load("dropin") is None or O.load("xdropin") is None,
                                 reason="tests/native/_build not built (make -C tests/native needs /root/reference)")]


def test_descriptor_distance_static_member():
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 256, (50, 32), dtype=np.uint8), rng.integers(0, 256, (50, 32), dtype=np.uint8)
    for x, y in zip(a, b):
        assert O.distance_ref(x, y, impl="dropin") == O.distance(x, y)
    assert O.distance_ref(np.zeros(32, np.uint8), np.full(32, 255, np.uint8), impl="dropin") == 256


@pytest.mark.parametrize("seed,window,check_ori", [(0, 100, True), (2, 100, False), (4, 10, True)])
def test_search_for_initialization_through_the_class(seed, window, check_ori):
    k1, d1, k2, d2 = _frame_pair(O, seed)
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
    a = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), prev, window, 0.9, check_ori)
    b = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), prev, window, 0.9, check_ori, impl="dropin")
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    a2 = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), a[2], window, 0.9, check_ori)
    b2 = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), b[2], window, 0.9, check_ori, impl="dropin")
    assert a2[0] == b2[0] and np.array_equal(a2[1], b2[1]) and np.array_equal(a2[2], b2[2])


@pytest.mark.parametrize("nmp,th,with_stereo,with_obs", [(3000, 3.0, False, False), (8000, 5.0, True, True)])
def test_search_by_projection_points_through_the_class(nmp, th, with_stereo, with_obs):
    """a13: SearchByProjection(Frame&, vector<MapPoint*>&, th) (src/ORBmatcher.cc:62-157)."""
    k, d, mp, mp_desc, rng = _projection_case(O, 7, nmp)
    n = len(k)
    sf = O.extractor("port").scale_tables()[0]
    ur = np.where(rng.random(n) < 0.6, k["x"] - rng.uniform(0, 12, n), -1).astype(np.float32) if with_stereo else None
    fmp0, fobs0, mobs = np.full(n, -1, np.int32), np.zeros(n, np.int32), np.ones(nmp, np.int32)
    if with_obs:
        held = rng.random(n) < 0.15
        fmp0[held] = 0
        fobs0[held] = rng.random(held.sum()) < 0.6
        mobs = (rng.random(nmp) < 0.8).astype(np.int32)
    a = O.search_by_projection_points(k, d, ur, (0, 1241, 0, 376), sf, mp, mp_desc, mobs, th, 0.8, fmp0, fobs0)
    b = O.search_by_projection_points(k, d, ur, (0, 1241, 0, 376), sf, mp, mp_desc, mobs, th, 0.8, fmp0, fobs0, impl="dropin")
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    assert a[0] > 50


@pytest.mark.parametrize("offset,th,mono,check_ori", [((0, 0, 0), 15.0, False, True), ((0, 0, 0.5), 15.0, False, True),
                                                      ((0, 0, -0.5), 7.0, False, True), ((0.5, 0, 0), 15.0, True, False)])
def test_search_by_projection_frame_through_the_class(offset, th, mono, check_ori):
    """a14: the per-frame tracking matcher SearchByProjection(Cur, Last, th, bMono, Calib) (src/ORBmatcher.cc:3448-3641),
    the call of Tracking::TrackWithMotionModel (src/Tracking.cc:1267)."""
    s = _rig_scene(O, 3, 1500, offset)
    sf = O.extractor("port").scale_tables()[0]
    n = s["n"]
    fmp0, fobs0 = np.full(n, -1, np.int32), np.zeros(n, np.int32)
    held = s["rng"].random(n) < 0.1
    fmp0[held] = 0
    fobs0[held] = s["rng"].random(held.sum()) < 0.5
    args = (s["cur_k"], s["cur_d"], s["ur"], s["cur_cam"], (0, 640, 0, 480), sf, CAM, s["Tcw"], s["Tlw"], s["last_k"], s["last_cam"],
            s["last_valid"], s["last_xyz"], s["last_desc"], s["last_obs"], CALIB, th, mono, check_ori, fmp0, fobs0)
    a = O.search_by_projection_frame(*args)
    b = O.search_by_projection_frame(*args, impl="dropin")
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    assert a[0] > 200


@pytest.mark.parametrize("th,orb_dist,check_ori", [(10.0, 100, True), (3.0, 64, True), (10.0, 100, False)])
def test_search_by_projection_keyframe_through_the_class(th, orb_dist, check_ori):
    """SearchByProjection(Cur, KeyFrame*, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:3809-3937); MapPoint::mfMaxDistance is
    read through the protected-member accessor of the drop-in."""
    from test_matcher_ref import _keyframe_case
    args = _keyframe_case(th, orb_dist, check_ori)
    a = O.search_by_projection_keyframe(*args)
    b = O.search_by_projection_keyframe(*args, impl="dropin")
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    assert a[0] > 100


@pytest.mark.parametrize("th,scale", [(10, 1.0), (4, 1.7)])
def test_search_by_projection_sim3_through_the_class(th, scale):
    """SearchByProjection(KeyFrame*, Scw, vpPoints, vLoopMPCams, vpMatched, th, Calib) (src/ORBmatcher.cc:566-752)."""
    from test_matcher_ref import _sim3_case
    args = _sim3_case(th, scale)
    a = O.search_by_projection_sim3(*args)
    b = O.search_by_projection_sim3(*args, impl="dropin")
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    assert a[0] > 50


def test_extractor_operator_call_through_the_class():
    """ORBextractor::operator() of the reference's class, GPU-backed (src/Frame.cc:397-403 calls it like this), with the
    optional mvImagePyramid mirror laid out as the reference's: ROI inside a parent with the 19-px reflected border."""
    from multi_orb_slam_b200.synth import textured
    port, drop = O.extractor("port"), O.extractor("xdropin")
    for seed, (w, h) in ((0, (640, 480)), (3, (752, 480)), (5, (640, 480))):  # the size change re-creates the device workspace
        img = textured(w, h, seed)
        kr, dr, _ = port.extract(img)
        kd, dd, _ = drop.extract(img)
        assert len(kd) == len(kr)
        for name in ("x", "y", "octave", "response", "size", "angle"):
            assert np.array_equal(kd[name], kr[name]), name
        assert np.array_equal(dd, dr)
        for level in (0, 3, 7):
            assert np.array_equal(drop.pyramid_level(level, 19), port.pyramid_level(level, 19)), level
    assert np.array_equal(drop.scale_tables()[0], port.scale_tables()[0])
