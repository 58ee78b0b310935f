"""Dev tool: randomised parity runs of the paths that changed in round 2 against the CPU oracle (seeds and shapes the test
suite does not hold).  usage: python tools/dev/fuzz_parity.py [seconds] [seed]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from multi_orb_slam_b200._lib import Camera
from multi_orb_slam_b200.extractor import ORBextractor
from multi_orb_slam_b200.matcher import Frame, ORBmatcher
from multi_orb_slam_b200.synth import RIG_CALIB, RIG_CAM, bow_scene, feature_vector, rig_scene, shifted_noisy, textured

O.build_oracle()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
t_end = time.time() + budget
stats = {"extract": 0, "sfi": 0, "bow": 0, "frame": 0}
SIZES = [(640, 480), (752, 480), (320, 240), (500, 400), (1024, 768), (333, 257), (1241, 376), (812, 612)]


def same_features(k, d, kr, dr):
    if len(k) != len(kr):
        return False
    return all(np.array_equal(k[f], kr[f]) for f in ("x", "y", "octave", "response", "size", "angle")) and np.array_equal(d, dr)


def image(w, h, seed, kind):
    if kind == 0:
        return textured(w, h, seed)
    r = np.random.default_rng(seed)
    if kind == 1:
        return r.integers(0, 256, (h, w), dtype=np.uint8)
    img = (128 + r.integers(-5, 6, (h, w))).astype(np.uint8)  # low texture + a few blobs
    for _ in range(6):
        y, x = int(r.integers(0, h - 40)), int(r.integers(0, w - 40))
        img[y:y + 30, x:x + 30] += np.uint8(r.integers(20, 90))
    return img


while time.time() < t_end:
    seed = int(rng.integers(0, 1 << 30))
    # --- extractor, the three FAST forms -------------------------------------------------------
    (w, h), nf, kind = SIZES[int(rng.integers(len(SIZES)))], int(rng.integers(200, 2200)), int(rng.integers(3))
    img = image(w, h, seed, kind)
    kr, dr, _ = O.extractor("port", nfeatures=nf).extract(img)
    for mode in ("cells", "bands", "split"):
        os.environ["ORB_B200_FAST"] = mode
        mb = int(rng.integers(1, 4))
        if os.environ.get("FUZZ_VERBOSE"):
            print(f"extract mode {mode} size {w}x{h} nf {nf} kind {kind} seed {seed} max_batch {mb}", flush=True)
        ex = ORBextractor(nf, 1.2, 8, 20, 7, image_size=(w, h), max_batch=mb)
        k, d = ex(img)
        if not same_features(k, d, kr, dr):
            raise SystemExit(f"EXTRACT MISMATCH mode {mode} size {w}x{h} nf {nf} kind {kind} seed {seed}")
        ex.close()
        stats["extract"] += 1
    os.environ["ORB_B200_FAST"] = "cells"
    # --- SearchForInitialization ------------------------------------------------------------------
    img1 = textured(640, 480, seed)
    img2 = shifted_noisy(img1, seed + 1)
    port = O.extractor("port")
    (k1, d1, _), (k2, d2, _) = port.extract(img1), port.extract(img2)
    window, ori = int(rng.choice([10, 30, 100, 300, 1000])), bool(rng.integers(2))
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32) + rng.normal(0, 3, (len(k1), 2)).astype(np.float32)
    rn, rm12, rprev = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), prev, window, 0.9, ori)
    gprev = prev.copy()
    gn, gm12 = ORBmatcher(0.9, ori).SearchForInitialization(Frame(k1, d1, 640, 480), Frame(k2, d2, 640, 480), gprev, window)
    if gn != rn or not np.array_equal(gm12, rm12) or not np.array_equal(gprev, rprev):
        raise SystemExit(f"SFI MISMATCH window {window} ori {ori} seed {seed}")
    stats["sfi"] += 1
    # --- SearchByBoW -------------------------------------------------------------------------------
    n1, n2, nodes = int(rng.integers(50, 2500)), int(rng.integers(50, 2500)), int(rng.choice([1, 3, 20, 100, 400, 1500]))
    sc = bow_scene(n1, n2, nodes, seed % 100000)
    node1 = np.where(rng.random(n1) < 0.05, -1, sc["node1"])
    node2 = np.where(rng.random(n2) < 0.05, -1, sc["node2"])
    fv1, fv2 = feature_vector(node1), feature_vector(node2)
    v1 = (rng.random(n1) < 0.8).astype(np.int32) if rng.integers(2) else None
    v2 = (rng.random(n2) < 0.9).astype(np.int32) if rng.integers(2) else None
    kf, ori = bool(rng.integers(2)), bool(rng.integers(2))
    got = ORBmatcher(0.7, ori).SearchByBoW(sc["d1"], sc["a1"], v1, fv1, sc["d2"], sc["a2"], v2, fv2, keyframe_pair=kf)
    ref = O.search_by_bow(sc["d1"], sc["a1"], v1, fv1, sc["d2"], sc["a2"], v2, fv2, 0.7, ori, 49 if kf else 50)
    if got[0] != ref[0] or not np.array_equal(got[1], ref[1]) or not np.array_equal(got[2], ref[2]):
        raise SystemExit(f"BOW MISMATCH n1 {n1} n2 {n2} nodes {nodes} kf {kf} ori {ori} seed {seed}")
    stats["bow"] += 1
    # --- SearchByProjection(CurrentFrame, LastFrame), two cameras, sometimes contended ------------------
    ports = {}
    s = rig_scene(lambda nfeat, im: ports.setdefault(nfeat, O.extractor("port", nfeatures=nfeat)).extract(im)[:2],
                  seed % 1000, int(rng.integers(200, 1800)), tuple(float(x) for x in rng.normal(0, 0.3, 3)))
    sf = port.scale_tables()[0]
    reps = int(rng.choice([1, 1, 2, 4]))
    rep = lambda a: np.repeat(a, reps, axis=0)
    lk, lc, lv, lx, ld = rep(s["last_k"]), rep(s["last_cam"]), rep(s["last_valid"]), rep(s["last_xyz"]), rep(s["last_desc"])
    lo = (rng.random(len(lk)) < 0.5).astype(s["last_obs"].dtype)
    n = s["n"]
    fmp0, fobs0 = np.full(n, -1, np.int32), np.zeros(n, np.int32)
    held = rng.random(n) < 0.1
    fmp0[held] = 0
    fobs0[held] = rng.random(int(held.sum())) < 0.5
    th, mono, ori = float(rng.choice([7.0, 15.0, 30.0])), bool(rng.integers(2)), bool(rng.integers(2))
    rn, rfmp = O.search_by_projection_frame(s["cur_k"], s["cur_d"], s["ur"], s["cur_cam"], (0, 640, 0, 480), sf, RIG_CAM, s["Tcw"],
                                            s["Tlw"], lk, lc, lv, lx, ld, lo, RIG_CALIB, th, mono, ori, fmp0, fobs0)
    F = Frame(s["cur_k"], s["cur_d"], 640, 480, mvScaleFactors=sf, mvuRight=s["ur"], mvpMapPoints=fmp0.copy(), mvpMapPointsObserved=fobs0)
    gn = ORBmatcher(0.9, ori).SearchByProjectionFrame(F, s["cur_cam"], Camera(*RIG_CAM), s["Tcw"], s["Tlw"], lk, lc, lv, lx, ld, lo,
                                                      RIG_CALIB, th, mono)
    if gn != rn or not np.array_equal(F.mvpMapPoints, rfmp):
        raise SystemExit(f"FRAME PROJECTION MISMATCH reps {reps} th {th} mono {mono} ori {ori} seed {seed}")
    stats["frame"] += 1
print("fuzz parity ok:", stats)
