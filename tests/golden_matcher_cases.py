"""Seeded matcher cases shared by tools/make_golden_matcher.py (which records what the reference's verbatim
ORBmatcher.cc build returns) and tests/test_golden.py (which checks the restatement against those records
where the verbatim build is not available).  Each case: run(O, impl) -> tuple of arrays / ints."""
import numpy as np


def _init(seed, window, check_ori):
    def run(O, impl):
        from test_gpu_matcher import _frame_pair
        k1, d1, k2, d2 = _frame_pair(O, seed)
        prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
        n, m12, newprev = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), prev, window, 0.9, check_ori, impl=impl)
        return n, m12, newprev
    return run


def _points(nmp, th):
    def run(O, impl):
        from test_gpu_matcher import _projection_case
        k, d, mp, mp_desc, rng = _projection_case(O, 7, nmp)
        n = len(k)
        sf = O.extractor("port").scale_tables()[0]
        ur = np.where(rng.random(n) < 0.6, k["x"] - rng.uniform(0, 12, n), -1).astype(np.float32)
        fmp0 = np.full(n, -1, np.int32)
        fobs0 = np.zeros(n, np.int32)
        held = rng.random(n) < 0.15
        fmp0[held] = 0
        fobs0[held] = rng.random(held.sum()) < 0.6
        mobs = (rng.random(nmp) < 0.8).astype(np.int32)
        return O.search_by_projection_points(k, d, ur, (0, 1241, 0, 376), sf, mp, mp_desc, mobs, th, 0.8, fmp0, fobs0, impl=impl)
    return run


def _frame(offset, th, mono):
    def run(O, impl):
        from test_gpu_matcher import CALIB, CAM, _rig_scene
        s = _rig_scene(O, 3, 1500, offset)
        sf = O.extractor("port").scale_tables()[0]
        n = s["n"]
        fmp0 = np.full(n, -1, np.int32)
        fobs0 = np.zeros(n, np.int32)
        held = s["rng"].random(n) < 0.1
        fmp0[held] = 0
        fobs0[held] = s["rng"].random(held.sum()) < 0.5
        return O.search_by_projection_frame(s["cur_k"], s["cur_d"], s["ur"], s["cur_cam"], (0, 640, 0, 480), sf, CAM, s["Tcw"], s["Tlw"],
                                            s["last_k"], s["last_cam"], s["last_valid"], s["last_xyz"], s["last_desc"], s["last_obs"],
                                            CALIB, th, mono, True, fmp0, fobs0, impl=impl)
    return run


def _bow(seed, n_nodes, variant):
    def run(O, impl):
        from multi_orb_slam_b200.synth import bow_scene, feature_vector
        n1, n2 = 900, 1000
        sc = bow_scene(n1, n2, n_nodes, seed)
        rng = np.random.default_rng(100 + seed)
        fv1 = feature_vector(np.where(rng.random(n1) < 0.05, -1, sc["node1"]))
        fv2 = feature_vector(np.where(sc["node2"] % 7 == 3, -1, sc["node2"]))
        v1 = (rng.random(n1) < 0.8).astype(np.int32)
        v2 = (rng.random(n2) < 0.9).astype(np.int32) if variant == 1 else None
        if impl == "ref":
            return O.search_by_bow_ref(variant, sc["d1"], sc["a1"], v1, fv1, sc["d2"], sc["a2"], v2, fv2, 0.7, True)
        return O.search_by_bow(sc["d1"], sc["a1"], v1, fv1, sc["d2"], sc["a2"], v2, fv2, 0.7, True, 50 if variant == 0 else 49)
    return run


CASES = {
    "init_s0_w100": _init(0, 100, True),
    "init_s1_w30": _init(1, 30, True),
    "points_3000_th1": _points(3000, 1.0),
    "frame_fwd": _frame((0, 0, 0.5), 15.0, False),
    "bow_kf_frame": _bow(0, 12, 0),
    "bow_kf_kf": _bow(1, 40, 1),
}
