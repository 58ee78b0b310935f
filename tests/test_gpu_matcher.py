"""GPU parity tests of the matcher kernels (through the C-ABI) against the flat-array CPU
restatement of src/ORBmatcher.cc (oracle/matcher_oracle.cc).  Everything here is integer /
index work: bit-exact."""
import numpy as np
import pytest

from multi_orb_slam_b200.synth import perturbed_descriptors, random_descriptors, shifted_noisy, textured

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def O(oracle_port):
    return oracle_port


@pytest.fixture(scope="module")
def M():
    from multi_orb_slam_b200.matcher import ORBmatcher
    return ORBmatcher(0.9, True)


def test_distance_known_answers(M):
    z, o = np.zeros(32, np.uint8), np.full(32, 255, np.uint8)
    assert M.DescriptorDistance(z, o) == 256
    assert M.DescriptorDistance(o, o) == 0
    for bit in (0, 7, 100, 255):
        a = z.copy()
        a[bit // 8] |= 1 << (bit % 8)
        assert M.DescriptorDistance(z, a) == 1


def test_distance_pairs_vs_oracle(M, O):
    a, b = random_descriptors(5000, 1), random_descriptors(5000, 2)
    got = M.distance_pairs(a, b)
    want = np.array([O.distance(a[i], b[i]) for i in range(0, 5000, 7)])
    assert np.array_equal(got[::7], want)
    assert np.array_equal(got, np.unpackbits(a ^ b, axis=1).sum(axis=1))


@pytest.mark.parametrize("nq,nt", [(1, 1), (1000, 1000), (777, 3001), (4096, 4096), (3, 70000), (257, 255),
                                   (8192, 8192), (16384, 16384), (32768, 32768)])  # the middle of configs[2]'s sweep
def test_bruteforce_vs_oracle(M, O, nq, nt):
    A = random_descriptors(max(nq, nt), 7)
    B, _ = perturbed_descriptors(A, 8)
    q, t = A[:nq], B[:nt]
    idx, d1, d2 = M.bruteforce(q, t, th_dist=50, ratio=0.9)
    ridx, rd1, rd2 = O.bruteforce(q, t, 0.9, 50)
    assert np.array_equal(d1, rd1) and np.array_equal(d2, rd2) and np.array_equal(idx, ridx)
    if nq >= 777:
        assert (idx >= 0).sum() > 0


def test_bruteforce_ties_take_lowest_index(M, O):
    t = np.repeat(random_descriptors(4, 3), 300, axis=0)  # many exact duplicates
    q = t[::150].copy()
    idx, d1, d2 = M.bruteforce(q, t, th_dist=50, ratio=2.0)
    ridx, rd1, rd2 = O.bruteforce(q, t, 2.0, 50)
    assert np.array_equal(idx, ridx) and np.array_equal(d1, rd1) and np.array_equal(d2, rd2)
    assert (d1 == 0).all() and (d2 == 0).all()


def test_bruteforce_empty_targets(M):
    idx, d1, d2 = M.bruteforce(random_descriptors(5, 1), np.zeros((0, 32), np.uint8))
    assert (idx == -1).all() and (d1 == 256).all() and (d2 == 256).all()


def test_bruteforce_full_size_properties(M):
    """64k x 64k (BASELINE.json config 3 upper end): size-independent properties instead of the
    oracle — every row of B is A[perm] with <=80 flips, so best distance equals the flip count
    and the match recovers the permutation wherever the ratio test accepts."""
    n = 65536
    A = random_descriptors(n, 7)
    B, perm = perturbed_descriptors(A, 8)
    idx, d1, d2 = M.bruteforce(B, A, th_dist=50, ratio=0.9)
    flips = np.unpackbits(B ^ A[perm], axis=1).sum(axis=1)
    assert np.array_equal(d1, np.minimum(d1, flips)) and (d1 <= flips).all()
    ok = idx >= 0
    assert ok.sum() > n // 3
    assert np.array_equal(idx[ok], perm[ok])
    assert (d1[ok] == flips[ok]).all() and (d1[ok] <= 50).all()
    assert (d1[ok].astype(np.float32) < np.float32(0.9) * d2[ok].astype(np.float32)).all()


def _frame_pair(O, seed):
    port = O.extractor("port")
    img1 = textured(640, 480, seed)
    img2 = shifted_noisy(img1, 1000 + seed)
    k1, d1, _ = port.extract(img1)
    k2, d2, _ = port.extract(img2)
    return k1, d1, k2, d2


@pytest.mark.parametrize("window,check_ori", [(100, True), (30, True), (100, False), (1000, True)])
def test_search_for_initialization_vs_oracle(M, O, window, check_ori):
    from multi_orb_slam_b200.matcher import Frame, ORBmatcher
    m = ORBmatcher(0.9, check_ori)
    for seed in (0, 1):
        k1, d1, k2, d2 = _frame_pair(O, seed)
        prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
        rn, rm12, rprev = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), prev, window, 0.9, check_ori)
        F1, F2 = Frame(k1, d1, 640, 480), Frame(k2, d2, 640, 480)
        gprev = prev.copy()
        gn, gm12 = m.SearchForInitialization(F1, F2, gprev, window)
        assert gn == rn and np.array_equal(gm12, rm12) and np.array_equal(gprev, rprev)
        assert rn > 20


def test_search_for_initialization_batch(M, O):
    pairs = [_frame_pair(O, s) for s in (3, 4, 5)]
    cap = max(max(len(p[0]), len(p[2])) for p in pairs)
    from multi_orb_slam_b200._lib import KP_DTYPE, Bounds
    P = len(pairs)
    k1 = np.zeros((P, cap), KP_DTYPE); k2 = np.zeros((P, cap), KP_DTYPE)
    d1 = np.zeros((P, cap, 32), np.uint8); d2 = np.zeros((P, cap, 32), np.uint8)
    prev = np.zeros((P, cap, 2), np.float32)
    n1 = np.zeros(P, np.int32); n2 = np.zeros(P, np.int32)
    for p, (a, da, b, db) in enumerate(pairs):
        n1[p], n2[p] = len(a), len(b)
        k1[p, : len(a)], d1[p, : len(a)], k2[p, : len(b)], d2[p, : len(b)] = a, da, b, db
        prev[p, : len(a), 0], prev[p, : len(a), 1] = a["x"], a["y"]
    nm, m12, newprev = M.search_for_initialization_batch(k1, d1, n1, k2, d2, n2, Bounds(0, 640, 0, 480), prev, 100)
    for p, (a, da, b, db) in enumerate(pairs):
        rn, rm12, rprev = O.search_for_initialization(a, da, b, db, (0, 640, 0, 480), prev[p, : len(a)], 100, 0.9, True)
        assert nm[p] == rn and np.array_equal(m12[p, : len(a)], rm12) and np.array_equal(newprev[p, : len(a)], rprev)


def _projection_case(O, seed, nmp):
    from multi_orb_slam_b200.synth import projection_case
    port = O.extractor("port", nfeatures=2000)
    k, d, _ = port.extract(textured(1241, 376, seed))
    mp, mpd, rng = projection_case(k, d, seed, nmp)
    return k, d, mp, mpd, rng


@pytest.mark.parametrize("nmp,th,with_stereo,with_obs", [(3000, 3.0, False, False), (20000, 3.0, False, False),
                                                          (5000, 1.0, True, True), (4000, 6.0, True, False)])
def test_search_by_projection_points_vs_oracle(M, O, nmp, th, with_stereo, with_obs):
    from multi_orb_slam_b200.matcher import Frame, MapPoints, ORBmatcher
    k, d, mp, mpd, rng = _projection_case(O, 2, nmp)
    sf = O.extractor("port").scale_tables()[0]
    n = len(k)
    ur = np.full(n, -1, np.float32)
    if with_stereo:
        ur = np.where(rng.random(n) < 0.6, k["x"] - rng.uniform(0, 12, n), -1).astype(np.float32)
    mp_obs = np.ones(nmp, np.int32)
    fmp0 = np.full(n, -1, np.int32)
    fobs0 = np.zeros(n, np.int32)
    if with_obs:
        mp_obs = (rng.random(nmp) < 0.8).astype(np.int32)
        held = rng.random(n) < 0.2
        fmp0[held] = rng.integers(0, nmp, held.sum())
        fobs0[held] = rng.random(held.sum()) < 0.7
    m = ORBmatcher(0.8, True)
    rn, rfmp = O.search_by_projection_points(k, d, ur, (0, 1241, 0, 376), sf, mp, mpd, mp_obs, th, 0.8, fmp0, fobs0)
    F = Frame(k, d, 1241, 376, mvScaleFactors=sf, mvuRight=ur, mvpMapPoints=fmp0.copy(), mvpMapPointsObserved=fobs0)
    gn = m.SearchByProjection(F, MapPoints(mp, mpd, mp_obs), th)
    assert gn == rn and np.array_equal(F.mvpMapPoints, rfmp)
    assert rn > nmp // 20


def test_search_by_projection_points_view_cos_boundary(M, O):
    """RadiusByViewingCos (src/ORBmatcher.cc:151-157) compares the float against the double literal 0.998:
    view_cos == float32(0.998) = 0.99800003 counts as greater (radius 2.5), float32 just below does not (4.0)."""
    from multi_orb_slam_b200.matcher import Frame, MapPoints, ORBmatcher
    k, d, mp, mpd, rng = _projection_case(O, 6, 4000)
    edge = np.float32(0.998)
    mp["view_cos"][::2] = edge
    mp["view_cos"][1::4] = np.nextafter(edge, np.float32(0))
    sf = O.extractor("port").scale_tables()[0]
    n, nmp = len(k), len(mp)
    ur = np.full(n, -1, np.float32)
    obs = np.ones(nmp, np.int32)
    fmp0, fobs0 = np.full(n, -1, np.int32), np.zeros(n, np.int32)
    rn, rfmp = O.search_by_projection_points(k, d, ur, (0, 1241, 0, 376), sf, mp, mpd, obs, 3.0, 0.8, fmp0, fobs0)
    F = Frame(k, d, 1241, 376, mvScaleFactors=sf, mvuRight=ur, mvpMapPoints=fmp0.copy(), mvpMapPointsObserved=fobs0)
    gn = ORBmatcher(0.8, True).SearchByProjection(F, MapPoints(mp, mpd, obs), 3.0)
    assert gn == rn and np.array_equal(F.mvpMapPoints, rfmp)
    # the boundary matters on this case: widening the edge points' window to 4.0 changes the oracle's answer
    mp2 = mp.copy()
    mp2["view_cos"][::2] = np.nextafter(edge, np.float32(0))
    rn2, rfmp2 = O.search_by_projection_points(k, d, ur, (0, 1241, 0, 376), sf, mp2, mpd, obs, 3.0, 0.8, fmp0, fobs0)
    assert not np.array_equal(rfmp2, rfmp)


def test_projection_searches_repeat_once_when_the_row_budget_overflows(M, O):
    """The single-call searches size their candidate rows from earlier calls (no host round trip between count and fill);
    a budget of one row per query overflows on the first call, which must repeat with the exact size and still equal
    the oracle — for the points overload (a13) and the tracking matcher (a14); later calls reuse the learnt size."""
    from multi_orb_slam_b200._lib import Camera
    from multi_orb_slam_b200.matcher import Frame, MapPoints, ORBmatcher
    k, d, mp, mpd, rng = _projection_case(O, 8, 5000)
    sf = O.extractor("port").scale_tables()[0]
    n, nmp = len(k), len(mp)
    ur, obs = np.full(n, -1, np.float32), np.ones(nmp, np.int32)
    fmp0, fobs0 = np.full(n, -1, np.int32), np.zeros(n, np.int32)
    rn, rfmp = O.search_by_projection_points(k, d, ur, (0, 1241, 0, 376), sf, mp, mpd, obs, 3.0, 0.8, fmp0, fobs0)
    m = ORBmatcher(0.8, True)
    m.debug_set_row_budget(1)
    for _ in range(2):  # first call overflows and repeats; the second one fits at once
        F = Frame(k, d, 1241, 376, mvScaleFactors=sf, mvuRight=ur, mvpMapPoints=fmp0.copy(), mvpMapPointsObserved=fobs0)
        assert m.SearchByProjection(F, MapPoints(mp, mpd, obs), 3.0) == rn and np.array_equal(F.mvpMapPoints, rfmp)
    s = _rig_scene(O, 3, 1500, (0, 0, 0.5))
    n = s["n"]
    fmp0, fobs0 = np.full(n, -1, np.int32), np.zeros(n, np.int32)
    rn, rfmp = O.search_by_projection_frame(s["cur_k"], s["cur_d"], s["ur"], s["cur_cam"], (0, 640, 0, 480), sf, CAM, s["Tcw"],
                                            s["Tlw"], s["last_k"], s["last_cam"], s["last_valid"], s["last_xyz"],
                                            s["last_desc"], s["last_obs"], CALIB, 15.0, False, True, fmp0, fobs0)
    m9 = ORBmatcher(0.9, True)
    m9.debug_set_row_budget(1)
    for _ in range(2):
        F = Frame(s["cur_k"], s["cur_d"], 640, 480, mvScaleFactors=sf, mvuRight=s["ur"], mvpMapPoints=fmp0.copy(),
                  mvpMapPointsObserved=fobs0)
        gn = m9.SearchByProjectionFrame(F, s["cur_cam"], Camera(*CAM), s["Tcw"], s["Tlw"], s["last_k"], s["last_cam"],
                                        s["last_valid"], s["last_xyz"], s["last_desc"], s["last_obs"], CALIB, 15.0, False)
        assert gn == rn and np.array_equal(F.mvpMapPoints, rfmp)


@pytest.mark.parametrize("reps,th", [(2, 3.0), (6, 6.0)])
def test_search_by_projection_points_contended(M, O, reps, th):
    """Every map point repeated `reps` times (same projection and descriptor, mixed Observations()): consecutive points
    compete for the same keypoints and for each other's second-best, so the order-dependent parts decide the result."""
    from multi_orb_slam_b200.matcher import Frame, MapPoints, ORBmatcher
    k, d, mp, mpd, rng = _projection_case(O, 4, 1500)
    mp, mpd = np.repeat(mp, reps), np.repeat(mpd, reps, axis=0)
    nmp = len(mp)
    sf = O.extractor("port").scale_tables()[0]
    n = len(k)
    ur = np.full(n, -1, np.float32)
    mp_obs = (rng.random(nmp) < 0.6).astype(np.int32)
    fmp0 = np.full(n, -1, np.int32)
    fobs0 = np.zeros(n, np.int32)
    held = rng.random(n) < 0.1
    fmp0[held] = rng.integers(0, nmp, held.sum())
    fobs0[held] = rng.random(held.sum()) < 0.5
    rn, rfmp = O.search_by_projection_points(k, d, ur, (0, 1241, 0, 376), sf, mp, mpd, mp_obs, th, 0.8, fmp0, fobs0)
    F = Frame(k, d, 1241, 376, mvScaleFactors=sf, mvuRight=ur, mvpMapPoints=fmp0.copy(), mvpMapPointsObserved=fobs0)
    gn = ORBmatcher(0.8, True).SearchByProjection(F, MapPoints(mp, mpd, mp_obs), th)
    assert gn == rn and np.array_equal(F.mvpMapPoints, rfmp)
    assert rn > 300


def test_bruteforce_batch_vs_oracle(O):
    import torch
    from multi_orb_slam_b200.matcher import ORBmatcher
    m = ORBmatcher(0.9, True)
    P, cap = 7, 1100
    rng = np.random.default_rng(12)
    A = random_descriptors(P * cap, 21).reshape(P, cap, 32)
    B = np.stack([perturbed_descriptors(A[p], 30 + p)[0] for p in range(P)])
    nq = rng.integers(0, cap + 1, P).astype(np.int32)
    nt = rng.integers(1, cap + 1, P).astype(np.int32)
    nq[0], nt[1] = cap, cap
    q, t = torch.from_numpy(B).cuda(), torch.from_numpy(A).cuda()
    idx, d1, d2 = (torch.full((P, cap), -7, dtype=torch.int32, device="cuda") for _ in range(3))
    m.bruteforce_batch_device(q, torch.from_numpy(nq).cuda(), t, torch.from_numpy(nt).cuda(), idx, d1, d2)
    m.sync()
    idx, d1, d2 = idx.cpu().numpy(), d1.cpu().numpy(), d2.cpu().numpy()
    for p in range(P):
        if nq[p] == 0:
            continue
        ridx, rd1, rd2 = O.bruteforce(B[p, : nq[p]], A[p, : nt[p]], 0.9, 50)
        assert np.array_equal(idx[p, : nq[p]], ridx) and np.array_equal(d1[p, : nq[p]], rd1) and np.array_equal(d2[p, : nq[p]], rd2)
        assert (idx[p, nq[p]:] == -7).all()


# ---- pose-based SearchByProjection overloads -------------------------------------------------
from multi_orb_slam_b200.synth import RIG_CALIB as CALIB, RIG_CAM as CAM, rig_rotation as _rot, rig_scene  # noqa: E402


def _rig_scene(O, seed, n_last, tlast_offset):
    """Two-camera current frame + last-frame map points (synth.rig_scene) on the ORACLE's features."""
    ports = {}

    def extract(nfeatures, image):
        port = ports.setdefault(nfeatures, O.extractor("port", nfeatures=nfeatures))
        return port.extract(image)[:2]

    return rig_scene(extract, seed, n_last, tlast_offset)


@pytest.mark.parametrize("offset,th,mono,check_ori", [((0, 0, 0), 15.0, False, True), ((0, 0, 0.5), 15.0, False, True),
                                                      ((0.4, 0, -0.5), 7.0, False, True), ((0, 0, 0.5), 30.0, True, False)])
def test_search_by_projection_frame_vs_oracle(O, offset, th, mono, check_ori):
    from multi_orb_slam_b200._lib import Camera
    from multi_orb_slam_b200.matcher import Frame, ORBmatcher
    s = _rig_scene(O, 3, 1500, offset)
    sf = O.extractor("port").scale_tables()[0]
    n = s["n"]
    fmp0 = np.full(n, -1, np.int32)
    fobs0 = np.zeros(n, np.int32)
    held = s["rng"].random(n) < 0.1
    fmp0[held] = 0
    fobs0[held] = s["rng"].random(held.sum()) < 0.5
    rn, rfmp = O.search_by_projection_frame(s["cur_k"], s["cur_d"], s["ur"], s["cur_cam"], (0, 640, 0, 480), sf, CAM, s["Tcw"],
                                            s["Tlw"], s["last_k"], s["last_cam"], s["last_valid"], s["last_xyz"],
                                            s["last_desc"], s["last_obs"], CALIB, th, mono, check_ori, fmp0, fobs0)
    F = Frame(s["cur_k"], s["cur_d"], 640, 480, mvScaleFactors=sf, mvuRight=s["ur"], mvpMapPoints=fmp0.copy(),
              mvpMapPointsObserved=fobs0)
    m = ORBmatcher(0.9, check_ori)
    gn = m.SearchByProjectionFrame(F, s["cur_cam"], Camera(*CAM), s["Tcw"], s["Tlw"], s["last_k"], s["last_cam"],
                                   s["last_valid"], s["last_xyz"], s["last_desc"], s["last_obs"], CALIB, th, mono)
    assert gn == rn and np.array_equal(F.mvpMapPoints, rfmp)
    assert rn > 200


@pytest.mark.parametrize("reps,th", [(2, 15.0), (5, 30.0)])
def test_search_by_projection_frame_contended(O, reps, th):
    """Every last-frame point repeated `reps` times (same projection, same descriptor, mixed Observations()):
    consecutive queries compete for the same keypoints, so the walk's order-dependent parts decide the result."""
    from multi_orb_slam_b200._lib import Camera
    from multi_orb_slam_b200.matcher import Frame, ORBmatcher
    s = _rig_scene(O, 5, 700, (0, 0, 0.3))
    sf = O.extractor("port").scale_tables()[0]
    n = s["n"]
    rep = lambda a: np.repeat(a, reps, axis=0)
    last_k, last_cam, last_valid = rep(s["last_k"]), rep(s["last_cam"]), rep(s["last_valid"])
    last_xyz, last_desc = rep(s["last_xyz"]), rep(s["last_desc"])
    last_obs = (s["rng"].random(len(last_k)) < 0.5).astype(s["last_obs"].dtype)
    fmp0 = np.full(n, -1, np.int32)
    fobs0 = np.zeros(n, np.int32)
    args = (CAM, s["Tcw"], s["Tlw"], last_k, last_cam, last_valid, last_xyz, last_desc, last_obs, CALIB, th, False)
    rn, rfmp = O.search_by_projection_frame(s["cur_k"], s["cur_d"], s["ur"], s["cur_cam"], (0, 640, 0, 480), sf, *args, True,
                                            fmp0, fobs0)
    F = Frame(s["cur_k"], s["cur_d"], 640, 480, mvScaleFactors=sf, mvuRight=s["ur"], mvpMapPoints=fmp0.copy(),
              mvpMapPointsObserved=fobs0)
    gn = ORBmatcher(0.9, True).SearchByProjectionFrame(F, s["cur_cam"], Camera(*CAM), *args[1:])
    assert gn == rn and np.array_equal(F.mvpMapPoints, rfmp)
    assert rn > 100


@pytest.mark.parametrize("th,orb_dist,check_ori", [(10.0, 100, True), (3.0, 64, True), (10.0, 100, False)])
def test_search_by_projection_keyframe_vs_oracle(O, th, orb_dist, check_ori):
    from multi_orb_slam_b200._lib import Camera
    from multi_orb_slam_b200.matcher import Frame, ORBmatcher
    s = _rig_scene(O, 9, 1200, (0, 0, 0))
    sel = s["cur_cam"] == 0  # this overload sees camera-1 (first camera) data only
    cur_k, cur_d = s["cur_k"][sel], s["cur_d"][sel]
    keep = s["last_cam"] == 0
    xyz, desc, ang = s["last_xyz"][keep], s["last_desc"][keep], s["last_k"]["angle"][keep]
    nk = len(xyz)
    rng = s["rng"]
    sf = O.extractor("port").scale_tables()[0]
    Ow = -s["Tcw"][:3, :3].T @ s["Tcw"][:3, 3]
    dist = np.linalg.norm(xyz - Ow, axis=1)
    max_d = (dist * rng.uniform(0.8, 4.0, nk)).astype(np.float32)  # mfMaxDistance
    kf_max, kf_min = (1.2 * max_d).astype(np.float32), (0.8 * max_d / 1.2 ** 7).astype(np.float32)
    valid = (rng.random(nk) < 0.9).astype(np.int32)
    fmp0 = np.full(len(cur_k), -1, np.int32)
    fmp0[rng.random(len(cur_k)) < 0.1] = 5
    log_sf = float(np.log(np.float32(1.2)))
    rn, rfmp = O.search_by_projection_keyframe(cur_k, cur_d, (0, 640, 0, 480), sf, log_sf, CAM, s["Tcw"], valid, xyz, kf_max,
                                               kf_min, max_d, ang, desc, th, orb_dist, check_ori, fmp0)
    F = Frame(cur_k, cur_d, 640, 480, mvScaleFactors=sf, mvpMapPoints=fmp0.copy())
    m = ORBmatcher(0.9, check_ori)
    gn = m.SearchByProjectionKeyFrame(F, Camera(*CAM), s["Tcw"], log_sf, valid, xyz, kf_max, kf_min, max_d, ang, desc, th, orb_dist)
    assert gn == rn and np.array_equal(F.mvpMapPoints, rfmp)
    assert rn > 100


@pytest.mark.parametrize("th,scale", [(10, 1.0), (4, 1.7)])
def test_search_by_projection_sim3_vs_oracle(O, th, scale):
    from multi_orb_slam_b200._lib import Camera
    from multi_orb_slam_b200.matcher import Frame, ORBmatcher
    s = _rig_scene(O, 13, 1800, (0, 0, 0))
    rng = s["rng"]
    n, nmp = s["n"], len(s["last_xyz"])
    sf = O.extractor("port").scale_tables()[0]
    Scw = s["Tcw"].astype(np.float64).copy()
    Scw[:3, :] *= scale                          # Sim3: s*R | s*t ; map points live in the scaled world
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
    rn, rout = O.search_by_projection_sim3(s["cur_k"], s["cur_d"], s["cur_cam"], (0, 640, 0, 480), sf, log_sf, CAM, Scw, CALIB,
                                           valid, xyz, normal, kf_max, kf_min, max_d, s["last_desc"], th, matched0)
    F = Frame(s["cur_k"], s["cur_d"], 640, 480, mvScaleFactors=sf, mvpMapPoints=matched0.copy())
    m = ORBmatcher(0.9, True)
    gn = m.SearchByProjectionSim3(F, s["cur_cam"], Camera(*CAM), log_sf, Scw, CALIB, valid, xyz, normal, kf_max, kf_min, max_d,
                                  s["last_desc"], th)
    assert gn == rn and np.array_equal(F.mvpMapPoints, rout)
    assert rn > 50


# ---- SearchByBoW (src/ORBmatcher.cc:206-388, 390-565, 996-1163, 1180-1363) ----------------------
@pytest.mark.parametrize("n1,n2,n_nodes,check_ori,kf_pair,with_valid", [
    (1000, 1100, 100, True, False, False),   # Frame variant, DBoW2 level-4 like node count
    (2000, 2000, 100, True, True, True),     # KeyFrame pair: strict threshold, both sides filtered
    (1500, 1500, 3, True, False, True),      # huge nodes: candidate rows stay in HBM (> 12288 entries)
    (300, 5000, 1000, False, False, False),  # sparse nodes, no orientation check
])
def test_search_by_bow_vs_oracle(O, n1, n2, n_nodes, check_ori, kf_pair, with_valid):
    from multi_orb_slam_b200.matcher import ORBmatcher
    from multi_orb_slam_b200.synth import bow_scene, feature_vector
    sc = bow_scene(n1, n2, n_nodes, 5 + n1)
    rng = np.random.default_rng(n2)
    node1 = np.where(rng.random(n1) < 0.03, -1, sc["node1"])
    node2 = np.where(sc["node2"] % 11 == 5, -1, sc["node2"])  # nodes missing on one side: lower_bound jumps
    fv1, fv2 = feature_vector(node1), feature_vector(node2)
    v1 = (rng.random(n1) < 0.8).astype(np.int32) if with_valid else None
    v2 = (rng.random(n2) < 0.9).astype(np.int32) if with_valid else None
    m = ORBmatcher(0.7, check_ori)
    nm, m12, m21 = m.SearchByBoW(sc["d1"], sc["a1"], v1, fv1, sc["d2"], sc["a2"], v2, fv2, keyframe_pair=kf_pair)
    rn, rm12, rm21 = O.search_by_bow(sc["d1"], sc["a1"], v1, fv1, sc["d2"], sc["a2"], v2, fv2, 0.7, check_ori,
                                     49 if kf_pair else 50)
    assert nm == rn and rn > 20
    assert np.array_equal(m12, rm12) and np.array_equal(m21, rm21)


def test_search_by_bow_edge_cases(O):
    from multi_orb_slam_b200.matcher import ORBmatcher
    from multi_orb_slam_b200.synth import bow_scene, feature_vector
    m = ORBmatcher(0.7, True)
    sc = bow_scene(50, 60, 4, 3)
    fv1, fv2 = feature_vector(sc["node1"]), feature_vector(sc["node2"] + 1000)  # no common node
    nm, m12, m21 = m.SearchByBoW(sc["d1"], sc["a1"], None, fv1, sc["d2"], sc["a2"], None, fv2)
    assert nm == 0 and (m12 == -1).all() and (m21 == -1).all()
    empty = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    nm, m12, m21 = m.SearchByBoW(sc["d1"], sc["a1"], None, empty, sc["d2"], sc["a2"], None, feature_vector(sc["node2"]))
    assert nm == 0 and (m12 == -1).all()
    # every side-1 feature invalid
    nm, m12, _ = m.SearchByBoW(sc["d1"], sc["a1"], np.zeros(50, np.int32), fv1, sc["d2"], sc["a2"], None,
                               feature_vector(sc["node2"]))
    assert nm == 0 and (m12 == -1).all()
    # identical sets in one node: ties resolved by vector order like the reference
    d = np.repeat(sc["d1"][:1], 8, axis=0)
    fv = feature_vector(np.zeros(8, np.int64))
    a = np.zeros(8, np.float32)
    got = m.SearchByBoW(d, a, None, fv, d, a, None, fv)
    ref = O.search_by_bow(d, a, None, fv, d, a, None, fv, 0.7, True, 50)
    assert got[0] == ref[0] and np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
    # bad index in a feature vector is rejected
    from multi_orb_slam_b200._lib import OrbError
    bad = (np.zeros(1, np.int32), np.array([0, 1], np.int32), np.array([99], np.int32))
    with pytest.raises(OrbError):
        m.SearchByBoW(sc["d1"], sc["a1"], None, bad, sc["d2"], sc["a2"], None, feature_vector(np.zeros(60, np.int64)))


# ---- SearchForTriangulation (src/ORBmatcher.cc:1364-1720) ---------------------------------------
@pytest.mark.parametrize("n1,n2,n_nodes,only_stereo,cam_enabled,check_ori", [
    (2000, 2200, 100, False, (True, True), True),
    (2000, 2000, 40, False, (True, False), True),     # vbCam disables camera 2
    (1500, 1500, 10, True, (True, True), False),      # bOnlyStereo
    (500, 6000, 2, False, (True, True), True),        # very large nodes
])
def test_search_for_triangulation_vs_oracle(O, n1, n2, n_nodes, only_stereo, cam_enabled, check_ori):
    from multi_orb_slam_b200.matcher import ORBmatcher
    from multi_orb_slam_b200.synth import feature_vector, triangulation_scene
    sc = triangulation_scene(n1, n2, n_nodes, 11 + n_nodes)
    node2 = np.where(sc["node2"] % 9 == 4, -1, sc["node2"])
    fv1, fv2 = feature_vector(sc["node1"]), feature_vector(node2)
    m = ORBmatcher(0.6, check_ori)
    nm, m12, pairs = m.SearchForTriangulation(sc["k1"], sc["d1"], sc["has_mp1"], sc["cam1"], sc["uright1"], fv1,
                                              sc["k2"], sc["d2"], sc["has_mp2"], sc["cam2"], sc["uright2"], fv2,
                                              sc["F12s"], sc["epipoles"], sc["scale_factors"], sc["level_sigma2"],
                                              bOnlyStereo=only_stereo, vbCam=cam_enabled)
    rn, rm12 = O.search_for_triangulation(sc, fv1, fv2, only_stereo, [int(v) for v in cam_enabled], check_ori)
    assert nm == rn and np.array_equal(m12, rm12)
    assert len(pairs) == nm and np.array_equal(pairs[:, 1], rm12[pairs[:, 0]])
    if not only_stereo:
        assert nm > 50
    # nothing matched where the reference forbids it
    assert not sc["has_mp1"][pairs[:, 0]].any() and not sc["has_mp2"][pairs[:, 1]].any()
    assert np.array_equal(sc["cam1"][pairs[:, 0]], sc["cam2"][pairs[:, 1]])


def test_search_for_triangulation_degenerate(O):
    from multi_orb_slam_b200.matcher import ORBmatcher
    from multi_orb_slam_b200.synth import feature_vector, triangulation_scene
    sc = triangulation_scene(60, 60, 3, 2)
    m = ORBmatcher(0.6, True)
    fv1 = feature_vector(sc["node1"])
    args = lambda fv2, F: (sc["k1"], sc["d1"], sc["has_mp1"], sc["cam1"], sc["uright1"], fv1, sc["k2"], sc["d2"], sc["has_mp2"],
                           sc["cam2"], sc["uright2"], fv2, F, sc["epipoles"], sc["scale_factors"], sc["level_sigma2"])
    nm, m12, pairs = m.SearchForTriangulation(*args(feature_vector(sc["node2"] + 50), sc["F12s"]))  # no common node
    assert nm == 0 and (m12 == -1).all() and pairs.shape == (0, 2)
    # zero fundamental matrix: den == 0 -> CheckDistEpipolarLine is false for every pair (:178-179)
    nm, m12, _ = m.SearchForTriangulation(*args(feature_vector(sc["node2"]), np.zeros((2, 3, 3), np.float32)))
    assert nm == 0 and (m12 == -1).all()


# ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:381-424) --------------------------
def test_compute_distinctive_descriptors_vs_oracle(M, O):
    from test_matcher_oracle import _distinctive_sets
    rng = np.random.default_rng(4)
    sizes = [1, 2, 0, 3, 160, 161, 400, 33] + [int(v) for v in rng.integers(1, 60, 3000)]
    desc, off = _distinctive_sets(1, sizes)
    got = M.ComputeDistinctiveDescriptors(desc, off)
    ref = O.compute_distinctive_descriptors(desc, off)
    assert np.array_equal(got, ref)
    assert got[2] == -1 and (got[np.array(sizes) > 0] >= 0).all()
    # all-identical set: every median is 0, the first descriptor wins
    same = np.repeat(random_descriptors(1, 3), 9, axis=0)
    assert M.ComputeDistinctiveDescriptors(same, [0, 9])[0] == 0


def test_search_by_bow_batch_vs_oracle(O):
    """A batch of independent (key frame, frame) pairs of different sizes in one call."""
    from multi_orb_slam_b200.matcher import ORBmatcher
    from multi_orb_slam_b200.synth import bow_scene, feature_vector
    m = ORBmatcher(0.7, True)
    pairs, refs = [], []
    rng = np.random.default_rng(6)
    for p, (n1, n2, nodes) in enumerate([(1000, 1100, 100), (50, 40, 5), (2000, 1500, 64), (0, 30, 4), (700, 700, 2), (300, 0, 9)]
                                        + [(int(a), int(b), 50) for a, b in rng.integers(200, 1500, (10, 2))]):
        sc = bow_scene(max(n1, 1), max(n2, 1), nodes, 40 + p)
        d1, a1, node1 = sc["d1"][:n1], sc["a1"][:n1], sc["node1"][:n1]
        d2, a2, node2 = sc["d2"][:n2], sc["a2"][:n2], sc["node2"][:n2]
        v1 = (rng.random(n1) < 0.85).astype(np.int32) if p % 2 else None
        v2 = (rng.random(n2) < 0.9).astype(np.int32) if p % 3 == 0 else None
        fv1, fv2 = feature_vector(node1), feature_vector(np.where(node2 % 13 == 1, -1, node2))
        pairs.append((d1, a1, v1, fv1, d2, a2, v2, fv2))
        refs.append(O.search_by_bow(d1, a1, v1, fv1, d2, a2, v2, fv2, 0.7, True, 50))
    got = m.SearchByBoW_batch(pairs)
    assert len(got) == len(refs)
    for p, ((nm, m12, m21), (rn, rm12, rm21)) in enumerate(zip(got, refs)):
        assert nm == rn and np.array_equal(m12, rm12) and np.array_equal(m21, rm21), f"pair {p}"
    assert sum(g[0] for g in got) > 500
    assert m.SearchByBoW_batch([]) == []


def test_search_for_triangulation_batch_vs_oracle(O):
    from multi_orb_slam_b200.matcher import ORBmatcher
    from multi_orb_slam_b200.synth import feature_vector, triangulation_scene
    m = ORBmatcher(0.6, True)
    scenes, refs = [], []
    for p, (n1, n2, nodes) in enumerate([(1500, 1600, 80), (40, 50, 3), (2000, 900, 30), (600, 2500, 4)] + [(800, 800, 50)] * 8):
        sc = triangulation_scene(n1, n2, nodes, 60 + p)
        sc["fv1"], sc["fv2"] = feature_vector(sc["node1"]), feature_vector(np.where(sc["node2"] % 7 == 3, -1, sc["node2"]))
        scenes.append(sc)
        refs.append(O.search_for_triangulation(sc, sc["fv1"], sc["fv2"], False, (1, 1), True))
    got = m.SearchForTriangulation_batch(scenes)
    for p, ((nm, m12), (rn, rm12)) in enumerate(zip(got, refs)):
        assert nm == rn and np.array_equal(m12, rm12), f"pair {p}"
    assert sum(g[0] for g in got) > 300
    assert m.SearchForTriangulation_batch([]) == []


# ---- Fuse(KeyFrame*, vpMapPoints, CalibMatrix, th) (src/ORBmatcher.cc:1986-2190) ----------------
@pytest.mark.parametrize("th,seed", [(3.0, 21), (6.0, 22)])
def test_fuse_vs_oracle(O, th, seed):
    from multi_orb_slam_b200._lib import Camera
    from multi_orb_slam_b200.matcher import Frame, ORBmatcher
    s = _rig_scene(O, seed, 2500, (0, 0, 0))
    rng = s["rng"]
    n, nmp = s["n"], len(s["last_xyz"])
    sf, _, sigma2, inv_sigma2 = O.extractor("port").scale_tables()
    Tcw = s["Tcw"]
    xyz = s["last_xyz"].astype(np.float64)
    R, t = Tcw[:3, :3].astype(np.float64), Tcw[:3, 3].astype(np.float64)
    Ow0 = -R.T @ t
    R12, t12 = CALIB[:3].astype(np.float64), CALIB[3].astype(np.float64)
    Ow1 = Ow0 + R.T @ t12  # centre of camera 2 in the world
    Ow = np.stack([Ow0, Ow1])
    dist = np.linalg.norm(xyz - Ow0, axis=1)
    normal = (xyz - Ow0) / dist[:, None] + rng.normal(0, 0.3, (nmp, 3))
    normal /= np.linalg.norm(normal, axis=1)[:, None]
    max_d = (dist * rng.uniform(0.8, 4.0, nmp)).astype(np.float32)
    kf_max, kf_min = (1.2 * max_d).astype(np.float32), (0.8 * max_d / 1.2 ** 7).astype(np.float32)
    valid = (rng.random(nmp) < 0.9).astype(np.int32)
    log_sf = float(np.log(np.float32(1.2)))
    rn, rbest = O.fuse(s["cur_k"], s["cur_d"], s["ur"], s["cur_cam"], (0, 640, 0, 480), sf, inv_sigma2, log_sf, CAM, Tcw, Ow, CALIB,
                       valid, xyz, normal, kf_max, kf_min, max_d, s["last_desc"], th)
    F = Frame(s["cur_k"], s["cur_d"], 640, 480, mvuRight=s["ur"], mvScaleFactors=sf)
    m = ORBmatcher(0.6, True)
    gn, gbest = m.Fuse(F, s["cur_cam"], Camera(*CAM), log_sf, inv_sigma2, Tcw, Ow, CALIB, valid, xyz, normal, kf_max, kf_min, max_d,
                       s["last_desc"], th)
    assert gn == rn and np.array_equal(gbest, rbest)
    assert rn > 100 and (rbest[valid == 0] == -1).all()
    # both chi-square branches were exercised (with and without a right coordinate)
    hit = rbest[rbest >= 0]
    assert (s["ur"][hit] >= 0).any() and (s["ur"][hit] < 0).any()
    # nothing to fuse
    gn, gbest = m.Fuse(F, s["cur_cam"], Camera(*CAM), log_sf, inv_sigma2, Tcw, Ow, CALIB, np.zeros(nmp, np.int32), xyz, normal,
                       kf_max, kf_min, max_d, s["last_desc"], th)
    assert gn == 0 and (gbest == -1).all()


@pytest.mark.parametrize("th,scale", [(4.0, 1.0), (4.0, 1.6)])
def test_fuse_sim3_vs_oracle(O, th, scale):
    """Fuse(KeyFrame*, Scw, ...) (src/ORBmatcher.cc:2211-2441): the loop-closing overload."""
    from multi_orb_slam_b200._lib import Camera
    from multi_orb_slam_b200.matcher import Frame, ORBmatcher
    s = _rig_scene(O, 31, 2500, (0, 0, 0))
    rng = s["rng"]
    nmp = len(s["last_xyz"])
    sf = O.extractor("port").scale_tables()[0]
    Scw = s["Tcw"].astype(np.float64).copy()
    Scw[:3, :] *= scale
    xyz = s["last_xyz"].astype(np.float64)
    Ow = -s["Tcw"][:3, :3].T.astype(np.float64) @ s["Tcw"][:3, 3].astype(np.float64)
    dist = np.linalg.norm(xyz - Ow, axis=1)
    normal = (xyz - Ow) / dist[:, None] + rng.normal(0, 0.3, (nmp, 3))
    normal /= np.linalg.norm(normal, axis=1)[:, None]
    max_d = (dist * rng.uniform(0.8, 4.0, nmp)).astype(np.float32)
    kf_max, kf_min = (1.2 * max_d).astype(np.float32), (0.8 * max_d / 1.2 ** 7).astype(np.float32)
    valid = (rng.random(nmp) < 0.9).astype(np.int32)
    log_sf = float(np.log(np.float32(1.2)))
    rn, rbest = O.fuse_sim3(s["cur_k"], s["cur_d"], s["cur_cam"], (0, 640, 0, 480), sf, log_sf, CAM, Scw, CALIB, valid, xyz, normal,
                            kf_max, kf_min, max_d, s["last_desc"], th)
    F = Frame(s["cur_k"], s["cur_d"], 640, 480, mvScaleFactors=sf)
    gn, gbest = ORBmatcher(0.6, True).FuseSim3(F, s["cur_cam"], Camera(*CAM), log_sf, Scw, CALIB, valid, xyz, normal, kf_max, kf_min,
                                                max_d, s["last_desc"], th)
    assert gn == rn and np.array_equal(gbest, rbest)
    assert rn > 100 and (rbest[:, 0] >= 0).any() and (rbest[:, 1] >= 0).any()


# ---- SearchBySim3 (src/ORBmatcher.cc:2814-3136) -------------------------------------------------
def _sim3_scene(O, s12, th):
    """Two two-camera key frames observing the same points; key frame 2 = key frame 1 moved by a Sim3."""
    from multi_orb_slam_b200.synth import KP_DTYPE
    rng = np.random.default_rng(int(s12 * 10 + th))
    fx, fy, cx, cy, mb, mbf = CAM
    port0, port1 = O.extractor("port", nfeatures=1000), O.extractor("port", nfeatures=500)
    kA0, dA0, _ = port0.extract(textured(640, 480, 41))
    kA1, dA1, _ = port1.extract(textured(640, 480, 141))
    k1, d1 = np.concatenate([kA0, kA1]), np.concatenate([dA0, dA1])
    cam1 = np.concatenate([np.zeros(len(kA0), np.int32), np.ones(len(kA1), np.int32)])
    n1 = len(k1)
    sf = port0.scale_tables()[0]
    # 3-D point of every key-frame-1 feature, in the frame of its own camera, then in the world (T1w = identity-ish)
    T1w = np.eye(4)
    T1w[:3, :3], T1w[:3, 3] = _rot(0.01, 0.02, -0.01), [0.1, 0.0, -0.05]
    R12c, t12c = CALIB[:3].astype(np.float64), CALIB[3].astype(np.float64)
    z = rng.uniform(1.5, 7.0, n1)
    Xown = np.stack([(k1["x"] - cx) / fx * z, (k1["y"] - cy) / fy * z, z], axis=1)
    Xc0 = np.where((cam1 == 1)[:, None], Xown @ R12c.T + t12c, Xown)   # camera-2 frame -> rig frame: R12 x + t12
    Xw = (Xc0 - T1w[:3, 3]) @ T1w[:3, :3]
    # Sim3 between the key frames: X1 = s12 * R12 * X2 + t12
    R12 = _rot(0.02, -0.015, 0.01)
    t12 = np.array([0.15, -0.03, 0.05])
    T2w = np.eye(4)
    R21 = R12.T
    T2w[:3, :3] = R21 @ T1w[:3, :3]
    T2w[:3, 3] = R21 @ (T1w[:3, 3] - t12) / 1.0
    # key frame 2: re-observations of random key-frame-1 features, projected with the inverse Sim3
    n2 = 1400
    src = rng.integers(0, n1, n2)
    X1 = Xw[src] @ T1w[:3, :3].T + T1w[:3, 3]
    X2 = ((X1 - t12) @ R12) / s12                                        # R12^T (X1 - t12) / s
    is1 = cam1[src] == 1
    R21c, t21c = np.linalg.inv(R12c), -np.linalg.inv(R12c) @ t12c
    X2own = np.where(is1[:, None], X2 @ R21c.T + t21c, X2)
    k2 = np.zeros(n2, KP_DTYPE)
    k2["x"] = fx * X2own[:, 0] / X2own[:, 2] + cx + rng.normal(0, 1.5, n2)
    k2["y"] = fy * X2own[:, 1] / X2own[:, 2] + cy + rng.normal(0, 1.5, n2)
    k2["octave"] = np.clip(k1["octave"][src] + rng.integers(-1, 2, n2), 0, 7)
    k2["angle"], k2["size"] = k1["angle"][src], 31
    bits = np.unpackbits(d1[src], axis=1)
    flips = rng.integers(0, 80, n2)
    bits ^= (np.argsort(np.argsort(rng.random((n2, 256)), axis=1), axis=1) < flips[:, None]).astype(np.uint8)
    d2 = np.packbits(bits, axis=1)
    cam2 = cam1[src].copy()
    # world points of key frame 2's map points (in ITS world: the reference uses each key frame's own pose)
    Xw2 = (X2 - T2w[:3, 3]) @ T2w[:3, :3]

    def points(xyz, Tw, n, desc, octave):
        Xc = xyz @ Tw[:3, :3].T + Tw[:3, 3]
        dist = np.linalg.norm(Xc, axis=1)
        max_d = (dist * 1.2 ** octave * rng.uniform(0.9, 1.1, n)).astype(np.float32)
        return dict(valid=(rng.random(n) < 0.85).astype(np.int32), xyz=xyz.astype(np.float32),
                    max_dist=(3.0 * max_d).astype(np.float32), min_dist=(0.2 * max_d / 1.2 ** 7).astype(np.float32), max_d=max_d,
                    desc=desc)

    mp1, mp2 = points(Xw, T1w, n1, d1, k1["octave"]), points(Xw2, T2w, n2, d2, k2["octave"])
    return dict(k1=k1, d1=d1, cam1=cam1, T1w=T1w, k2=k2, d2=d2, cam2=cam2, T2w=T2w, sf=sf, R12=R12, t12=t12, mp1=mp1, mp2=mp2,
                log_sf=float(np.log(np.float32(1.2))))


@pytest.mark.parametrize("s12,th", [(1.0, 7.5), (1.3, 7.5), (0.8, 3.0)])
def test_search_by_sim3_vs_oracle(O, s12, th):
    from multi_orb_slam_b200._lib import Camera
    from multi_orb_slam_b200.matcher import Frame, ORBmatcher
    c = _sim3_scene(O, s12, th)
    k1, d1, cam1, T1w, k2, d2, cam2, T2w, sf, R12, t12, mp1, mp2, log_sf = (c[k] for k in (
        "k1", "d1", "cam1", "T1w", "k2", "d2", "cam2", "T2w", "sf", "R12", "t12", "mp1", "mp2", "log_sf"))
    rn, rm12 = O.search_by_sim3(k1, d1, cam1, T1w, k2, d2, cam2, T2w, (0, 640, 0, 480), sf, log_sf, CAM, s12, R12, t12, CALIB,
                                mp1, mp2, th)
    F1 = Frame(k1, d1, 640, 480, mvScaleFactors=sf)
    F2 = Frame(k2, d2, 640, 480, mvScaleFactors=sf)
    gn, gm12 = ORBmatcher(0.75, True).SearchBySim3(F1, cam1, T1w, F2, cam2, T2w, Camera(*CAM), log_sf, s12, R12, t12, CALIB, mp1, mp2, th)
    assert gn == rn and np.array_equal(gm12, rm12)
    if s12 == 1.0:
        assert rn > 30


# ---- DBoW2 vocabulary transform (Frame::ComputeBoW) ----------------------------------------------
@pytest.mark.parametrize("k,L,levelsup,n", [(10, 4, 2, 2000), (10, 3, 4, 500), (4, 6, 4, 1500), (10, 5, 4, 3000)])
def test_compute_bow_vs_oracle(O, k, L, levelsup, n):
    from multi_orb_slam_b200.matcher import ORBmatcher
    from multi_orb_slam_b200.synth import random_vocabulary
    voc = random_vocabulary(k, L, 10 * k + L)
    rng = np.random.default_rng(n)
    leaves = np.nonzero(voc["word_id"] >= 0)[0]
    base = voc["node_desc"][rng.choice(leaves, n)]
    desc = np.packbits(np.unpackbits(base, axis=1) ^ (rng.random((n, 256)) < 0.05).astype(np.uint8), axis=1)
    m = ORBmatcher(0.7, True)
    m.set_vocabulary(voc["child_start"], voc["child_ids"], voc["node_desc"], voc["word_id"], voc["node_weight"], voc["L"])
    got = m.ComputeBoW(desc, levelsup)
    ref = O.bow_transform(voc, desc, levelsup)
    for key in ("word", "node", "weight"):
        assert np.array_equal(got[key], ref[key]), key
    assert np.array_equal(got["bow"][0], ref["bow"][0]) and np.array_equal(got["bow"][1], ref["bow"][1])  # doubles, bit for bit
    for a, b in zip(got["featvec"], ref["featvec"]):
        assert np.array_equal(a, b)
    # the feature vector feeds SearchByBoW: the frame against itself matches (almost) every feature to itself
    ang = np.zeros(n, np.float32)
    nm, m12, _ = m.SearchByBoW(desc, ang, None, got["featvec"], desc, ang, None, got["featvec"])
    rn, rm12, _ = O.search_by_bow(desc, ang, None, ref["featvec"], desc, ang, None, ref["featvec"], 0.7, True, 50)
    assert nm == rn and np.array_equal(m12, rm12)


def test_compute_bow_requires_vocabulary():
    from multi_orb_slam_b200._lib import OrbError
    from multi_orb_slam_b200.matcher import ORBmatcher
    with pytest.raises(OrbError):
        ORBmatcher(0.7, True).ComputeBoW(np.zeros((3, 32), np.uint8))
