"""Pins the matcher restatement (oracle/matcher_oracle.cc, the checker of every GPU matcher test) against the
reference's OWN src/ORBmatcher.cc, compiled verbatim from /root/reference into oracle/_ref/libmatcher_ref.so
(oracle/Makefile; OpenCV-API shim + Frame/KeyFrame/MapPoint stand-ins in oracle/shim_matcher).  Same seeded
scenes as tests/test_gpu_matcher.py.  Skipped where the verbatim build is absent."""
import numpy as np
import pytest

import oracle_lib as O
from test_gpu_matcher import CALIB, CAM, _frame_pair, _projection_case, _rig_scene

pytestmark = pytest.mark.skipif(O.load("mref") is None, reason="oracle/_ref/libmatcher_ref.so not built (no /root/reference)")


def test_descriptor_distance():
    rng = np.random.default_rng(0)
    a, b = rng.integers(0, 256, (200, 32), dtype=np.uint8), rng.integers(0, 256, (200, 32), dtype=np.uint8)
    for x, y in zip(a, b):
        assert O.distance_ref(x, y) == O.distance(x, y)
    assert O.distance_ref(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


@pytest.mark.parametrize("seed,window,check_ori", [(0, 100, True), (1, 30, True), (2, 100, False), (3, 1000, True), (4, 10, True)])
def test_search_for_initialization(seed, window, check_ori):
    k1, d1, k2, d2 = _frame_pair(O, seed)
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
    a = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), prev, window, 0.9, check_ori)
    b = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), prev, window, 0.9, check_ori, impl="ref")
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[0] > 20 or window == 10
    # second round with the updated vbPrevMatched, like Tracking::MonocularInitialization re-entering
    a2 = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), a[2], window, 0.9, check_ori)
    b2 = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), b[2], window, 0.9, check_ori, impl="ref")
    assert a2[0] == b2[0] and np.array_equal(a2[1], b2[1]) and np.array_equal(a2[2], b2[2])


@pytest.mark.parametrize("nmp,th,with_stereo,with_obs", [(3000, 3.0, False, False), (3000, 1.0, True, True), (8000, 5.0, True, True)])
def test_search_by_projection_points(nmp, th, with_stereo, with_obs):
    k, d, mp, mp_desc, rng = _projection_case(O, 7, nmp)
    n = len(k)
    sf = O.extractor("port").scale_tables()[0]
    ur = np.where(rng.random(n) < 0.6, k["x"] - rng.uniform(0, 12, n), -1).astype(np.float32) if with_stereo else None
    fmp0 = np.full(n, -1, np.int32)
    fobs0 = np.zeros(n, np.int32)
    mobs = np.ones(nmp, np.int32)
    if with_obs:
        held = rng.random(n) < 0.15
        fmp0[held] = 0
        fobs0[held] = rng.random(held.sum()) < 0.6
        mobs = (rng.random(nmp) < 0.8).astype(np.int32)
    a = O.search_by_projection_points(k, d, ur, (0, 1241, 0, 376), sf, mp, mp_desc, mobs, th, 0.8, fmp0, fobs0)
    b = O.search_by_projection_points(k, d, ur, (0, 1241, 0, 376), sf, mp, mp_desc, mobs, th, 0.8, fmp0, fobs0, impl="ref")
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    assert a[0] > 50


@pytest.mark.parametrize("offset,th,mono,check_ori", [((0, 0, 0), 15.0, False, True), ((0, 0, 0.5), 15.0, False, True),
                                                      ((0, 0, -0.5), 7.0, False, True), ((0.5, 0, 0), 15.0, True, False)])
def test_search_by_projection_frame(offset, th, mono, check_ori):
    s = _rig_scene(O, 3, 1500, offset)
    sf = O.extractor("port").scale_tables()[0]
    n = s["n"]
    fmp0 = np.full(n, -1, np.int32)
    fobs0 = np.zeros(n, np.int32)
    held = s["rng"].random(n) < 0.1
    fmp0[held] = 0
    fobs0[held] = s["rng"].random(held.sum()) < 0.5
    args = (s["cur_k"], s["cur_d"], s["ur"], s["cur_cam"], (0, 640, 0, 480), sf, CAM, s["Tcw"], s["Tlw"], s["last_k"], s["last_cam"],
            s["last_valid"], s["last_xyz"], s["last_desc"], s["last_obs"], CALIB, th, mono, check_ori, fmp0, fobs0)
    a = O.search_by_projection_frame(*args)
    b = O.search_by_projection_frame(*args, impl="ref")
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    assert a[0] > 100


def _keyframe_case(th, orb_dist, check_ori):
    s = _rig_scene(O, 9, 1200, (0, 0, 0))
    sel = s["cur_cam"] == 0
    cur_k, cur_d = s["cur_k"][sel], s["cur_d"][sel]
    keep = s["last_cam"] == 0
    xyz, desc, ang = s["last_xyz"][keep], s["last_desc"][keep], s["last_k"]["angle"][keep]
    nk = len(xyz)
    rng = s["rng"]
    sf = O.extractor("port").scale_tables()[0]
    Ow = -s["Tcw"][:3, :3].T @ s["Tcw"][:3, 3]
    dist = np.linalg.norm(xyz - Ow, axis=1)
    max_d = (dist * rng.uniform(0.8, 4.0, nk)).astype(np.float32)
    kf_max, kf_min = (1.2 * max_d).astype(np.float32), (0.8 * max_d / 1.2 ** 7).astype(np.float32)
    valid = (rng.random(nk) < 0.9).astype(np.int32)
    fmp0 = np.full(len(cur_k), -1, np.int32)
    fmp0[rng.random(len(cur_k)) < 0.1] = 5
    log_sf = float(np.log(np.float32(1.2)))
    return (cur_k, cur_d, (0, 640, 0, 480), sf, log_sf, CAM, s["Tcw"], valid, xyz, kf_max, kf_min, max_d, ang, desc, th, orb_dist,
            check_ori, fmp0)


@pytest.mark.parametrize("th,orb_dist,check_ori", [(10.0, 100, True), (3.0, 64, True), (10.0, 100, False)])
def test_search_by_projection_keyframe(th, orb_dist, check_ori):
    args = _keyframe_case(th, orb_dist, check_ori)
    a = O.search_by_projection_keyframe(*args)
    b = O.search_by_projection_keyframe(*args, impl="ref")
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    assert a[0] > 100


def _sim3_case(th, scale):
    s = _rig_scene(O, 13, 1800, (0, 0, 0))
    rng = s["rng"]
    n, nmp = s["n"], len(s["last_xyz"])
    sf = O.extractor("port").scale_tables()[0]
    Scw = s["Tcw"].astype(np.float64).copy()
    Scw[:3, :] *= scale
    xyz = s["last_xyz"].astype(np.float64)
    Ow = -s["Tcw"][:3, :3].T.astype(np.float64) @ s["Tcw"][:3, 3].astype(np.float64)
    PO = xyz - Ow
    dist = np.linalg.norm(PO, axis=1)
    normal = PO / dist[:, None] + rng.normal(0, 0.3, (nmp, 3))
    normal /= np.linalg.norm(normal, axis=1)[:, None]
    max_d = (dist * rng.uniform(0.8, 4.0, nmp)).astype(np.float32)
    kf_max, kf_min = (1.2 * max_d).astype(np.float32), (0.8 * max_d / 1.2 ** 7).astype(np.float32)
    valid = (rng.random(nmp) < 0.9).astype(np.int32)
    matched0 = np.full(n, -1, np.int32)
    matched0[rng.random(n) < 0.1] = 3
    log_sf = float(np.log(np.float32(1.2)))
    return (s["cur_k"], s["cur_d"], s["cur_cam"], (0, 640, 0, 480), sf, log_sf, CAM, Scw, CALIB, valid, xyz, normal, kf_max, kf_min,
            max_d, s["last_desc"], th, matched0)


@pytest.mark.parametrize("th,scale", [(10, 1.0), (4, 1.7)])
def test_search_by_projection_sim3(th, scale):
    args = _sim3_case(th, scale)
    a = O.search_by_projection_sim3(*args)
    b = O.search_by_projection_sim3(*args, impl="ref")
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    assert a[0] > 50


@pytest.mark.parametrize("seed,n_nodes,check_ori,variant", [(0, 12, True, 0), (1, 40, True, 1), (2, 5, False, 0), (3, 200, True, 1),
                                                           (4, 100, True, 0)])
def test_search_by_bow(seed, n_nodes, check_ori, variant):
    from multi_orb_slam_b200.synth import bow_scene, feature_vector
    n1, n2 = 900, 1000
    sc = bow_scene(n1, n2, n_nodes, seed)
    rng = np.random.default_rng(100 + seed)
    node1 = np.where(rng.random(n1) < 0.05, -1, sc["node1"])
    node2 = np.where(sc["node2"] % 7 == 3, -1, sc["node2"])
    fv1, fv2 = feature_vector(node1), feature_vector(node2)
    v1 = (rng.random(n1) < 0.8).astype(np.int32)
    v2 = (rng.random(n2) < 0.9).astype(np.int32) if variant == 1 else None
    a = O.search_by_bow(sc["d1"], sc["a1"], v1, fv1, sc["d2"], sc["a2"], v2, fv2, 0.7, check_ori, 50 if variant == 0 else 49)
    b = O.search_by_bow_ref(variant, sc["d1"], sc["a1"], v1, fv1, sc["d2"], sc["a2"], v2, fv2, 0.7, check_ori)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2])
    assert a[0] > 20


@pytest.mark.parametrize("seed,n_nodes,only_stereo,cam_enabled,check_ori", [(0, 30, False, (1, 1), True), (1, 8, False, (1, 0), True),
                                                                          (2, 15, True, (1, 1), False)])
def test_search_for_triangulation(seed, n_nodes, only_stereo, cam_enabled, check_ori):
    """The reference recomputes F12 and the epipoles from the poses (:1376-1449); the flat implementations take them as
    inputs, so the verbatim run also returns them (same expressions, same shim) for the restatement to consume."""
    from multi_orb_slam_b200.synth import feature_vector, triangulation_scene
    sc = triangulation_scene(900, 1000, n_nodes, seed)
    fv1, fv2 = feature_vector(sc["node1"]), feature_vector(np.where(sc["node2"] % 5 == 2, -1, sc["node2"]))
    # poses with key frame 2 at the origin: R12 = R1w, t12 = t1w per camera (X1 = R12 X2 + t12)
    fx, fy, cx, cy = 520.0, 520.0, 320.0, 240.0
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    T1w, Cw1 = np.zeros((2, 4, 4)), np.zeros((2, 3))
    for c in range(2):
        E = np.linalg.inv(K.T) @ np.zeros((3, 3))  # placeholder to keep K in scope
        F = sc["F12s"][c].astype(np.float64)
        # recover [t12]x R12 = K^T F K and split it with the generator's known rotation convention
        M = K.T @ F @ K
        U, S, Vt = np.linalg.svd(M)
        W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
        cands = [(U @ W @ Vt, U[:, 2]), (U @ W.T @ Vt, U[:, 2])]
        R = min((r * np.sign(np.linalg.det(r)) for r, _ in cands), key=lambda r: np.linalg.norm(r - np.eye(3)))
        tx = M @ R.T
        t = np.array([tx[2, 1], tx[0, 2], tx[1, 0]])
        T1w[c, :3, :3], T1w[c, :3, 3], T1w[c, 3, 3] = R, t, 1
        Cw1[c] = -R.T @ t
    T2w = np.stack([np.eye(4), np.eye(4)])
    cam = (fx, fy, cx, cy, 0.08, 40.0)
    rn, rm12, F12s, epi = O.search_for_triangulation_ref(sc, fv1, fv2, T1w, T2w, Cw1, cam, only_stereo, cam_enabled, check_ori)
    sc2 = dict(sc, F12s=F12s, epipoles=epi)
    pn, pm12 = O.search_for_triangulation(sc2, fv1, fv2, only_stereo, cam_enabled, check_ori)
    assert pn == rn and np.array_equal(pm12, rm12)
    if not only_stereo:
        assert rn > 20
    # the recomputed fundamental matrices are the generator's up to scale
    for c in range(2):
        a, b = F12s[c].astype(np.float64).ravel(), sc["F12s"][c].astype(np.float64).ravel()
        assert abs(abs(a @ b) / (np.linalg.norm(a) * np.linalg.norm(b)) - 1) < 1e-3


def _fuse_scene(seed):
    s = _rig_scene(O, seed, 2500, (0, 0, 0))
    rng = s["rng"]
    nmp = len(s["last_xyz"])
    Tcw = s["Tcw"]
    xyz = s["last_xyz"].astype(np.float64)
    R, t = Tcw[:3, :3].astype(np.float64), Tcw[:3, 3].astype(np.float64)
    Ow0 = -R.T @ t
    Ow = np.stack([Ow0, Ow0 + R.T @ CALIB[3].astype(np.float64)])
    dist = np.linalg.norm(xyz - Ow0, axis=1)
    normal = (xyz - Ow0) / dist[:, None] + rng.normal(0, 0.3, (nmp, 3))
    normal /= np.linalg.norm(normal, axis=1)[:, None]
    max_d = (dist * rng.uniform(0.8, 4.0, nmp)).astype(np.float32)
    return dict(s=s, xyz=xyz, Ow=Ow, normal=normal, max_d=max_d, kf_max=(1.2 * max_d).astype(np.float32),
                kf_min=(0.8 * max_d / 1.2 ** 7).astype(np.float32), valid=(rng.random(nmp) < 0.9).astype(np.int32),
                held=rng.integers(0, 3, s["n"]).astype(np.int32) * (rng.random(s["n"]) < 0.4), log_sf=float(np.log(np.float32(1.2))))


@pytest.mark.parametrize("th,seed", [(3.0, 21), (6.0, 22)])
def test_fuse(th, seed):
    f = _fuse_scene(seed)
    s = f["s"]
    sf, _, _, inv_sigma2 = O.extractor("port").scale_tables()
    a = O.fuse(s["cur_k"], s["cur_d"], s["ur"], s["cur_cam"], (0, 640, 0, 480), sf, inv_sigma2, f["log_sf"], CAM, s["Tcw"], f["Ow"],
               CALIB, f["valid"], f["xyz"], f["normal"], f["kf_max"], f["kf_min"], f["max_d"], s["last_desc"], th)
    b = O.fuse_ref(0, s["cur_k"], s["cur_d"], s["ur"], s["cur_cam"], f["held"], (0, 640, 0, 480), sf, inv_sigma2, f["log_sf"], CAM,
                   s["Tcw"], f["Ow"], CALIB, f["valid"], f["xyz"], f["normal"], f["kf_max"], f["kf_min"], f["max_d"], s["last_desc"], th)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    assert a[0] > 100


@pytest.mark.parametrize("th,scale", [(4.0, 1.0), (4.0, 1.6)])
def test_fuse_sim3(th, scale):
    f = _fuse_scene(31)
    s = f["s"]
    sf = O.extractor("port").scale_tables()[0]
    Scw = s["Tcw"].astype(np.float64).copy()
    Scw[:3, :] *= scale
    a = O.fuse_sim3(s["cur_k"], s["cur_d"], s["cur_cam"], (0, 640, 0, 480), sf, f["log_sf"], CAM, Scw, CALIB, f["valid"], f["xyz"],
                    f["normal"], f["kf_max"], f["kf_min"], f["max_d"], s["last_desc"], th)
    b = O.fuse_ref(1, s["cur_k"], s["cur_d"], None, s["cur_cam"], f["held"], (0, 640, 0, 480), sf, None, f["log_sf"], CAM, Scw,
                   np.zeros(6), CALIB, f["valid"], f["xyz"], f["normal"], f["kf_max"], f["kf_min"], f["max_d"], s["last_desc"], th)
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    assert a[0] > 100


@pytest.mark.parametrize("s12,th", [(1.0, 7.5), (1.3, 7.5), (0.8, 3.0)])
def test_search_by_sim3(s12, th):
    from test_gpu_matcher import _sim3_scene
    c = _sim3_scene(O, s12, th)
    args = (c["k1"], c["d1"], c["cam1"], c["T1w"], c["k2"], c["d2"], c["cam2"], c["T2w"], (0, 640, 0, 480), c["sf"], c["log_sf"], CAM,
            s12, c["R12"], c["t12"], CALIB, c["mp1"], c["mp2"], th)
    a = O.search_by_sim3(*args)
    b = O.search_by_sim3(*args, impl="ref")
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    if s12 == 1.0:
        assert a[0] > 30


# ---- the camera-1-only twins: the two-camera restatements fed with camera-1 data only ----------
@pytest.fixture
def cam1_twins():
    lib = O.load("mref")
    lib.omr_set_cam1(1)
    yield
    lib.omr_set_cam1(0)


def _cam0_only(s):
    """Camera-1 subset of a _rig_scene (the reference numbers camera-1 features first)."""
    sel = s["cur_cam"] == 0
    return s["cur_k"][sel], s["cur_d"][sel], np.zeros(int(sel.sum()), np.int32)


@pytest.mark.parametrize("th,scale", [(10, 1.0), (4, 1.7)])
def test_search_by_projection_sim3_cam1(cam1_twins, th, scale):
    s = _rig_scene(O, 13, 1800, (0, 0, 0))
    k, d, cam = _cam0_only(s)
    rng = s["rng"]
    n, nmp = len(k), len(s["last_xyz"])
    sf = O.extractor("port").scale_tables()[0]
    Scw = s["Tcw"].astype(np.float64).copy()
    Scw[:3, :] *= scale
    xyz = s["last_xyz"].astype(np.float64)
    Ow = -s["Tcw"][:3, :3].T.astype(np.float64) @ s["Tcw"][:3, 3].astype(np.float64)
    PO = xyz - Ow
    dist = np.linalg.norm(PO, axis=1)
    normal = PO / dist[:, None] + rng.normal(0, 0.3, (nmp, 3))
    normal /= np.linalg.norm(normal, axis=1)[:, None]
    max_d = (dist * rng.uniform(0.8, 4.0, nmp)).astype(np.float32)
    kf_max, kf_min = (1.2 * max_d).astype(np.float32), (0.8 * max_d / 1.2 ** 7).astype(np.float32)
    valid = (rng.random(nmp) < 0.9).astype(np.int32)
    matched0 = np.full(n, -1, np.int32)
    matched0[rng.random(n) < 0.1] = 3
    args = (k, d, cam, (0, 640, 0, 480), sf, float(np.log(np.float32(1.2))), CAM, Scw, CALIB, valid, xyz, normal, kf_max, kf_min, max_d,
            s["last_desc"], th, matched0)
    a = O.search_by_projection_sim3(*args)
    b = O.search_by_projection_sim3(*args, impl="ref")
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[0] > 30


@pytest.mark.parametrize("seed,n_nodes,variant", [(0, 12, 0), (1, 40, 1), (4, 100, 0)])
def test_search_by_bow_cam1(cam1_twins, seed, n_nodes, variant):
    from multi_orb_slam_b200.synth import bow_scene, feature_vector
    n1, n2 = 900, 1000
    sc = bow_scene(n1, n2, n_nodes, seed)
    rng = np.random.default_rng(100 + seed)
    fv1, fv2 = feature_vector(sc["node1"]), feature_vector(np.where(sc["node2"] % 7 == 3, -1, sc["node2"]))
    v1 = (rng.random(n1) < 0.8).astype(np.int32)
    v2 = (rng.random(n2) < 0.9).astype(np.int32) if variant == 1 else None
    a = O.search_by_bow(sc["d1"], sc["a1"], v1, fv1, sc["d2"], sc["a2"], v2, fv2, 0.7, True, 50 if variant == 0 else 49)
    b = O.search_by_bow_ref(variant, sc["d1"], sc["a1"], v1, fv1, sc["d2"], sc["a2"], v2, fv2, 0.7, True)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[0] > 20


def test_fuse_cam1(cam1_twins):
    f = _fuse_scene(31)
    s = f["s"]
    k, d, cam = _cam0_only(s)
    sf = O.extractor("port").scale_tables()[0]
    Scw = s["Tcw"].astype(np.float64).copy()
    Scw[:3, :] *= 1.3
    held = f["held"][s["cur_cam"] == 0]
    a = O.fuse_sim3(k, d, cam, (0, 640, 0, 480), sf, f["log_sf"], CAM, Scw, CALIB, f["valid"], f["xyz"], f["normal"], f["kf_max"],
                    f["kf_min"], f["max_d"], s["last_desc"], 4.0)
    b = O.fuse_ref(1, k, d, None, cam, held, (0, 640, 0, 480), sf, None, f["log_sf"], CAM, Scw, np.zeros(6), CALIB, f["valid"], f["xyz"],
                   f["normal"], f["kf_max"], f["kf_min"], f["max_d"], s["last_desc"], 4.0)
    assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[0] > 50 and (a[1][:, 1] == -1).all()


@pytest.mark.parametrize("s12,th", [(1.0, 7.5), (1.3, 7.5)])
def test_search_by_sim3_cam1(cam1_twins, s12, th):
    from test_gpu_matcher import _sim3_scene
    c = _sim3_scene(O, s12, th)
    s1, s2 = c["cam1"] == 0, c["cam2"] == 0
    sub = lambda mp, m: {k: v[m] for k, v in mp.items()}
    args = (c["k1"][s1], c["d1"][s1], np.zeros(int(s1.sum()), np.int32), c["T1w"], c["k2"][s2], c["d2"][s2], np.zeros(int(s2.sum()), np.int32),
            c["T2w"], (0, 640, 0, 480), c["sf"], c["log_sf"], CAM, s12, c["R12"], c["t12"], CALIB, sub(c["mp1"], s1), sub(c["mp2"], s2), th)
    a = O.search_by_sim3(*args)
    b = O.search_by_sim3(*args, impl="ref")
    assert a[0] == b[0] and np.array_equal(a[1], b[1])
    if s12 == 1.0:
        assert a[0] > 20


@pytest.mark.parametrize("k,L,levelsup,n", [(10, 4, 2, 1500), (10, 3, 4, 400), (4, 6, 4, 1000), (10, 5, 4, 2000)])
def test_dbow2_transform(tmp_path, k, L, levelsup, n):
    """Frame::ComputeBoW: the reference's DBoW2 (TemplatedVocabulary.h, FORB.cpp, BowVector.cpp, FeatureVector.cpp compiled
    verbatim) loads the synthetic vocabulary from an ORBvoc-format text file and transforms the same descriptors."""
    from multi_orb_slam_b200.synth import random_vocabulary
    voc = random_vocabulary(k, L, 10 * k + L)
    rng = np.random.default_rng(n)
    leaves = np.nonzero(voc["word_id"] >= 0)[0]
    base = voc["node_desc"][rng.choice(leaves, n)]
    desc = np.packbits(np.unpackbits(base, axis=1) ^ (rng.random((n, 256)) < 0.05).astype(np.uint8), axis=1)
    path = tmp_path / "voc.txt"
    O.write_vocabulary_text(voc, path, k)
    b = O.bow_transform_ref(path, desc, levelsup)
    a = O.bow_transform(voc, desc, levelsup)
    assert b["n_words"] == int((voc["word_id"] >= 0).sum())
    assert np.array_equal(a["bow"][0], b["bow"][0]) and np.array_equal(a["bow"][1], b["bow"][1])  # doubles, bit for bit
    for x, y in zip(a["featvec"], b["featvec"]):
        assert np.array_equal(x, y)
    assert len(a["bow"][0]) > 50
