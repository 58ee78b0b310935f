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


# ---- ComputeStereoMatches (src/Frame.cc:782-956) ------------------------------------------------------------------
def _stereo_oracle(O, seed, nfeatures=1000, W=640, H=480, **kw):
    from multi_orb_slam_b200.synth import stereo_pair
    left, right = stereo_pair(W, H, seed, **kw)
    out = []
    for img in (left, right):
        ex = O.extractor("port", nfeatures=nfeatures)
        k, d, _ = ex.extract(img)
        out.append((k, d, [ex.pyramid_level(l) for l in range(8)], ex.scale_tables()))
    return left, right, out


def test_stereo_matches_recover_the_band_disparities(oracle_port):
    O = oracle_port
    disp = (4, 9, 17, 30)
    left, right, ((kl, dl, pl, tl), (kr, dr, pr, _)) = _stereo_oracle(O, 3, disparities=disp)
    mbf, fx = 40.0, 500.0
    ur, z = O.compute_stereo_matches(kl, dl, kr, dr, pl, pr, tl[0], tl[1], mbf, mbf / fx)
    ok = ur >= 0
    assert ok.sum() > 0.4 * len(kl)
    assert ((ur < 0) == (z < 0)).all()
    band = np.minimum((kl["y"] * len(disp) / 480).astype(int), len(disp) - 1)
    want = np.asarray(disp, np.float32)[band]
    err = np.abs((kl["x"] - ur)[ok] - want[ok])
    # away from the band seams the sub-pixel disparity is the band's shift
    inner = np.abs(kl["y"][ok] - np.rint(kl["y"][ok] / 120) * 120) > 12 * tl[0][kl["octave"][ok]]
    assert np.median(err[inner]) < 0.25 and (err[inner] < 1.5 * tl[0][kl["octave"][ok]][inner]).mean() > 0.97
    assert np.allclose(z[ok], mbf / (kl["x"] - ur)[ok], rtol=1e-5)


def test_stereo_matches_python_transliteration(oracle_port):
    """Independent pure-Python statement of :782-956 on a small pair (300 features)."""
    O = oracle_port
    left, right, ((kl, dl, pl, tl), (kr, dr, pr, _)) = _stereo_oracle(O, 5, nfeatures=300, W=400, H=300, disparities=(6, 13))
    mbf, mb = np.float32(35.0), np.float32(0.09)
    got_u, got_z = O.compute_stereo_matches(kl, dl, kr, dr, pl, pr, tl[0], tl[1], mbf, mb)
    f32 = np.float32
    sf, isf = tl[0], tl[1]
    rows = [[] for _ in range(300)]
    for i, k in enumerate(kr):
        r = f32(2.0) * sf[k["octave"]]
        for y in range(int(np.floor(k["y"] - r)), int(np.ceil(k["y"] + r)) + 1):
            rows[y].append(i)
    maxD = mbf / mb
    u_out, z_out, lst = np.full(len(kl), -1, f32), np.full(len(kl), -1, f32), []
    rnd = lambda v: int(np.floor(v + f32(0.5)))  # positive values
    for i, k in enumerate(kl):
        cand = rows[int(k["y"])]
        best, bi = 100, 0
        for j in cand:
            if abs(int(kr[j]["octave"]) - int(k["octave"])) > 1 or not (k["x"] - maxD <= kr[j]["x"] <= k["x"]):
                continue
            d = int(np.unpackbits(dl[i] ^ dr[j]).sum())
            if d < best:
                best, bi = d, j
        if best >= 75:
            continue
        s = isf[k["octave"]]
        xl, yl, xr = rnd(k["x"] * s), rnd(k["y"] * s), rnd(kr[bi]["x"] * s)
        L, R = pl[k["octave"]].astype(np.int32), pr[k["octave"]].astype(np.int32)
        if xr < 0 or xr + 11 >= R.shape[1] - 38:
            continue
        a = L[19 + yl - 5:19 + yl + 6, 19 + xl - 5:19 + xl + 6] - L[19 + yl, 19 + xl]
        dists = []
        for inc in range(-5, 6):
            b = R[19 + yl - 5:19 + yl + 6, 19 + xr + inc - 5:19 + xr + inc + 6] - R[19 + yl, 19 + xr + inc]
            dists.append(f32(np.abs(a - b).sum()))
        inc = int(np.argmin(dists)) - 5
        if inc in (-5, 5):
            continue
        d1, d2, d3 = dists[inc + 4], dists[inc + 5], dists[inc + 6]
        with np.errstate(all="ignore"):
            delta = (d1 - d3) / (f32(2.0) * (d1 + d3 - f32(2.0) * d2))
        if delta < -1 or delta > 1:
            continue
        bu = sf[k["octave"]] * (f32(xr) + f32(inc) + delta)
        disp = k["x"] - bu
        if disp >= 0 and disp < maxD:
            if disp <= 0:
                disp, bu = f32(0.01), f32(np.float64(k["x"]) - 0.01)
            z_out[i], u_out[i] = mbf / disp, bu
            lst.append((int(dists[inc + 5]), i))
    lst.sort()
    th = f32(1.5) * f32(1.4) * f32(lst[len(lst) // 2][0])
    for sad, i in lst:
        if not f32(sad) < th:
            u_out[i] = z_out[i] = -1
    assert (u_out >= 0).sum() > 100
    assert np.array_equal(got_u, u_out) and np.array_equal(got_z, z_out)
