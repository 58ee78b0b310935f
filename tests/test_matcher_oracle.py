"""Known-answer tests for the matcher restatement (oracle/matcher_oracle.cc).  The reference ships
no tests or vectors for these functions; the primary pin is the verbatim build of the
reference's ORBmatcher.cc (tests/test_matcher_ref.py).  Here:
(1) hand-checkable cases and (2) an independent pure-Python transliteration of
src/ORBmatcher.cc:868-983 / src/Frame.cc:510-566,632-642 written from the reference text, run on
small random cases."""
import math

import numpy as np
import pytest

from multi_orb_slam_b200.synth import random_descriptors


@pytest.fixture(scope="module")
def O(oracle_port):
    return oracle_port


def test_distance_known_answers(O):
    z, o = np.zeros(32, np.uint8), np.full(32, 255, np.uint8)
    assert O.distance(z, o) == 256 and O.distance(o, o) == 0 and O.distance(z, z) == 0
    for bit in (0, 9, 128, 255):
        a = z.copy()
        a[bit // 8] ^= 1 << (bit % 8)
        assert O.distance(z, a) == 1
    a, b = random_descriptors(50, 1), random_descriptors(50, 2)
    for i in range(50):
        assert O.distance(a[i], b[i]) == int(np.unpackbits(a[i] ^ b[i]).sum())


def test_three_maxima(O):
    assert O.three_maxima([0] * 30) == (-1, -1, -1)
    c = [0] * 30
    c[3], c[7], c[9] = 10, 5, 2
    assert O.three_maxima(c) == (3, 7, 9)
    c[9] = 0            # max3 (0) < 0.1*max1 -> third dropped
    assert O.three_maxima(c) == (3, 7, -1)
    c[7] = 0            # max2 < 0.1*max1 -> second and third dropped
    assert O.three_maxima(c) == (3, -1, -1)
    c = [0] * 30
    c[2] = c[5] = c[8] = 4  # ties: first bin wins (strict >)
    assert O.three_maxima(c) == (2, 5, 8)


def test_bruteforce_known_answers(O):
    t = random_descriptors(64, 3)
    q = t[[5, 9]].copy()
    q[1, 0] ^= 0b111  # three flips
    idx, d1, d2 = O.bruteforce(q, t, 0.9, 50)
    assert list(idx) == [5, 9] and list(d1) == [0, 3] and (d2 > 60).all()
    dup = np.concatenate([t, t[5:6]])  # exact duplicate of target 5 later in the list: ratio test fails, best = first
    idx, d1, d2 = O.bruteforce(q[:1], dup, 0.9, 50)
    assert idx[0] == -1 and d1[0] == 0 and d2[0] == 0


def _py_grid(kx, ky, bounds):
    minx, maxx, miny, maxy = bounds
    inv_w = np.float32(64) / np.float32(maxx - minx)
    inv_h = np.float32(48) / np.float32(maxy - miny)
    cells = {}
    for i, (x, y) in enumerate(zip(kx, ky)):
        fx = np.float32(np.float32(x - np.float32(minx)) * inv_w)
        fy = np.float32(np.float32(y - np.float32(miny)) * inv_h)
        px = int(math.floor(abs(fx) + 0.5) * (1 if fx >= 0 else -1))  # C round(): half away from zero
        py = int(math.floor(abs(fy) + 0.5) * (1 if fy >= 0 else -1))
        if 0 <= px < 64 and 0 <= py < 48:
            cells.setdefault((px, py), []).append(i)
    return cells, inv_w, inv_h


def _py_area(cells, inv_w, inv_h, kx, ky, koct, bounds, x, y, r, lo, hi):
    minx, _, miny, _ = bounds
    f = np.float32
    cx0 = max(0, int(math.floor(f(f(f(x) - f(minx)) - f(r)) * inv_w)))
    cx1 = min(63, int(math.ceil(f(f(f(x) - f(minx)) + f(r)) * inv_w)))
    cy0 = max(0, int(math.floor(f(f(f(y) - f(miny)) - f(r)) * inv_h)))
    cy1 = min(47, int(math.ceil(f(f(f(y) - f(miny)) + f(r)) * inv_h)))
    if cx0 >= 64 or cx1 < 0 or cy0 >= 48 or cy1 < 0:
        return []
    check = lo > 0 or hi >= 0
    out = []
    for ix in range(cx0, cx1 + 1):
        for iy in range(cy0, cy1 + 1):
            for i in cells.get((ix, iy), []):
                if check and (koct[i] < lo or (hi >= 0 and koct[i] > hi)):
                    continue
                if abs(f(kx[i]) - f(x)) < r and abs(f(ky[i]) - f(y)) < r:
                    out.append(i)
    return out


def test_features_in_area_vs_python(O):
    rng = np.random.default_rng(4)
    n = 400
    kx = rng.uniform(-5, 645, n).astype(np.float32)
    ky = rng.uniform(-5, 485, n).astype(np.float32)
    koct = rng.integers(0, 8, n).astype(np.int32)
    bounds = (0.0, 640.0, 0.0, 480.0)
    cells, iw, ih = _py_grid(kx, ky, bounds)
    for _ in range(60):
        x, y = float(rng.uniform(-50, 700)), float(rng.uniform(-50, 530))
        r = float(rng.choice([3.0, 15.5, 100.0]))
        lo, hi = [(-1, -1), (0, 0), (2, 3), (1, -1)][int(rng.integers(0, 4))]
        got = O.features_in_area(kx, ky, koct, bounds, x, y, r, lo, hi)
        want = _py_area(cells, iw, ih, kx, ky, koct, bounds, np.float32(x), np.float32(y), np.float32(r), lo, hi)
        assert list(got) == want


def _py_search_init(k1, d1, k2, d2, bounds, prev, window, ratio, check_ori):
    n1, n2 = len(k1), len(k2)
    m12 = [-1] * n1
    m21 = [-1] * n2
    md = [2**31 - 1] * n2
    hist = [[] for _ in range(30)]
    cells, iw, ih = _py_grid(k2["x"], k2["y"], bounds)
    nm = 0
    dist = lambda a, b: int(np.unpackbits(a ^ b).sum())
    for i1 in range(n1):
        if k1["octave"][i1] > 0:
            continue
        cand = _py_area(cells, iw, ih, k2["x"], k2["y"], k2["octave"], bounds, prev[i1, 0], prev[i1, 1], np.float32(window), 0, 0)
        if not cand:
            continue
        best, best2, bi = 2**31 - 1, 2**31 - 1, -1
        for i2 in cand:
            dd = dist(d1[i1], d2[i2])
            if md[i2] <= dd:
                continue
            if dd < best:
                best2, best, bi = best, dd, i2
            elif dd < best2:
                best2 = dd
        if best <= 50 and np.float32(best) < np.float32(best2) * np.float32(ratio):
            if m21[bi] >= 0:
                m12[m21[bi]] = -1
                nm -= 1
            m12[i1], m21[bi], md[bi] = bi, i1, best
            nm += 1
            if check_ori:
                rot = np.float32(k1["angle"][i1]) - np.float32(k2["angle"][bi])
                if rot < 0:
                    rot = np.float32(rot + np.float32(360.0))
                v = np.float32(rot * np.float32(1.0 / 30))
                b = int(math.floor(v + 0.5))
                hist[0 if b == 30 else b].append(i1)
    if check_ori:
        import oracle_lib
        keep = set(oracle_lib.three_maxima([len(h) for h in hist]))
        for b in range(30):
            if b in keep:
                continue
            for i1 in hist[b]:
                if m12[i1] >= 0:
                    m12[i1] = -1
                    nm -= 1
    return nm, m12


@pytest.mark.parametrize("seed,window,check_ori", [(0, 40, True), (1, 100, True), (2, 15, False), (3, 1000, True)])
def test_search_for_initialization_vs_python(O, seed, window, check_ori):
    from oracle_lib import KP_DTYPE
    rng = np.random.default_rng(seed)
    n1, n2 = 120, 140
    k1 = np.zeros(n1, KP_DTYPE)
    k1["x"], k1["y"] = rng.uniform(0, 640, n1), rng.uniform(0, 480, n1)
    k1["octave"] = rng.integers(0, 3, n1)
    k1["angle"] = rng.uniform(0, 360, n1)
    d1 = random_descriptors(n1, seed + 10)
    src = rng.integers(0, n1, n2)
    k2 = np.zeros(n2, KP_DTYPE)
    k2["x"] = k1["x"][src] + rng.normal(0, 6, n2)
    k2["y"] = k1["y"][src] + rng.normal(0, 6, n2)
    k2["octave"] = np.where(rng.random(n2) < 0.8, 0, 1)
    k2["angle"] = (k1["angle"][src] + rng.normal(0, 20, n2)) % 360
    bits = np.unpackbits(d1[src], axis=1)
    flips = rng.integers(0, 70, n2)
    bits ^= (np.argsort(np.argsort(rng.random((n2, 256)), axis=1), axis=1) < flips[:, None]).astype(np.uint8)
    d2 = np.packbits(bits, axis=1)
    prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
    bounds = (0.0, 640.0, 0.0, 480.0)
    rn, rm12, rprev = O.search_for_initialization(k1, d1, k2, d2, bounds, prev, window, 0.9, check_ori)
    pn, pm12 = _py_search_init(k1, d1, k2, d2, bounds, prev, window, 0.9, check_ori)
    assert rn == pn and list(rm12) == pm12
    assert rn == int((rm12 >= 0).sum())
    for i1 in np.nonzero(rm12 >= 0)[0]:
        assert rprev[i1, 0] == k2["x"][rm12[i1]] and rprev[i1, 1] == k2["y"][rm12[i1]]


# ---- SearchByBoW (src/ORBmatcher.cc:206-388, 390-565, 996-1163, 1180-1363) ----------------------
def _py_search_by_bow(d1, a1, valid1, fv1, d2, a2, valid2, fv2, ratio, check_ori, max_dist):
    """Transliteration over real std::map-like containers (dict + sorted keys + bisect = lower_bound)."""
    import bisect
    map1 = {int(n): [int(i) for i in fv1[2][fv1[1][k]:fv1[1][k + 1]]] for k, n in enumerate(fv1[0])}
    map2 = {int(n): [int(i) for i in fv2[2][fv2[1][k]:fv2[1][k + 1]]] for k, n in enumerate(fv2[0])}
    keys1, keys2 = sorted(map1), sorted(map2)
    m12, m21 = [-1] * len(d1), [-1] * len(d2)
    hist = [[] for _ in range(30)]
    dist = lambda x, y: int(np.unpackbits(x ^ y).sum())
    nm, ia, ib = 0, 0, 0
    while ia < len(keys1) and ib < len(keys2):
        if keys1[ia] == keys2[ib]:
            for idx1 in map1[keys1[ia]]:
                if valid1 is not None and not valid1[idx1]:
                    continue
                best1, bi, best2 = 256, -1, 256
                for idx2 in map2[keys2[ib]]:
                    if m21[idx2] >= 0 or (valid2 is not None and not valid2[idx2]):
                        continue
                    dd = dist(d1[idx1], d2[idx2])
                    if dd < best1:
                        best2, best1, bi = best1, dd, idx2
                    elif dd < best2:
                        best2 = dd
                if best1 <= max_dist and np.float32(best1) < np.float32(ratio) * np.float32(best2):
                    m12[idx1], m21[bi] = bi, idx1
                    if check_ori:
                        rot = np.float32(a1[idx1]) - np.float32(a2[bi])
                        if rot < 0:
                            rot = np.float32(rot + np.float32(360.0))
                        b = int(math.floor(np.float32(rot * np.float32(1.0 / 30)) + 0.5))
                        hist[0 if b == 30 else b].append(idx1)
                    nm += 1
            ia, ib = ia + 1, ib + 1
        elif keys1[ia] < keys2[ib]:
            ia = bisect.bisect_left(keys1, keys2[ib])
        else:
            ib = bisect.bisect_left(keys2, keys1[ia])
    if check_ori:
        import oracle_lib
        keep = set(oracle_lib.three_maxima([len(h) for h in hist]))
        for b in range(30):
            if b not in keep:
                for idx1 in hist[b]:
                    m21[m12[idx1]] = -1
                    m12[idx1] = -1
                    nm -= 1
    return nm, m12, m21


@pytest.mark.parametrize("seed,n_nodes,check_ori,max_dist,with_valid", [(0, 12, True, 50, False), (1, 40, True, 49, True),
                                                                       (2, 5, False, 50, True), (3, 200, True, 50, False)])
def test_search_by_bow_vs_python(O, seed, n_nodes, check_ori, max_dist, with_valid):
    from multi_orb_slam_b200.synth import bow_scene, feature_vector
    sc = bow_scene(150, 170, n_nodes, seed)
    rng = np.random.default_rng(100 + seed)
    # leave some features out of the vectors and make the two node sets differ (lower_bound jumps)
    node1 = np.where(rng.random(150) < 0.05, -1, sc["node1"])
    node2 = np.where(rng.random(170) < 0.05, -1, sc["node2"])
    node2 = np.where(node2 % 7 == 3, -1, node2)
    fv1, fv2 = feature_vector(node1), feature_vector(node2)
    v1 = (rng.random(150) < 0.8).astype(np.int32) if with_valid else None
    v2 = (rng.random(170) < 0.9).astype(np.int32) if with_valid else None
    rn, rm12, rm21 = O.search_by_bow(sc["d1"], sc["a1"], v1, fv1, sc["d2"], sc["a2"], v2, fv2, 0.7, check_ori, max_dist)
    pn, pm12, pm21 = _py_search_by_bow(sc["d1"], sc["a1"], v1, fv1, sc["d2"], sc["a2"], v2, fv2, 0.7, check_ori, max_dist)
    assert rn == pn and list(rm12) == pm12 and list(rm21) == pm21
    assert rn == int((rm12 >= 0).sum()) == int((rm21 >= 0).sum()) and rn > 10
    for i1 in np.nonzero(rm12 >= 0)[0]:
        assert rm21[rm12[i1]] == i1


def test_search_by_bow_degenerate(O):
    from multi_orb_slam_b200.synth import bow_scene, feature_vector
    sc = bow_scene(20, 20, 3, 9)
    fv1, fv2 = feature_vector(sc["node1"]), feature_vector(sc["node2"] + 100)  # no common node
    rn, rm12, rm21 = O.search_by_bow(sc["d1"], sc["a1"], None, fv1, sc["d2"], sc["a2"], None, fv2)
    assert rn == 0 and (rm12 == -1).all() and (rm21 == -1).all()
    empty = (np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32))
    rn, rm12, _ = O.search_by_bow(sc["d1"], sc["a1"], None, empty, sc["d2"], sc["a2"], None, fv2)
    assert rn == 0 and (rm12 == -1).all()


# ---- SearchForTriangulation (src/ORBmatcher.cc:1364-1720, CheckDistEpipolarLine :167-184) -------
def _py_search_for_triangulation(sc, fv1, fv2, only_stereo, cam_enabled, check_ori):
    import bisect
    f32 = np.float32
    map1 = {int(n): [int(i) for i in fv1[2][fv1[1][k]:fv1[1][k + 1]]] for k, n in enumerate(fv1[0])}
    map2 = {int(n): [int(i) for i in fv2[2][fv2[1][k]:fv2[1][k + 1]]] for k, n in enumerate(fv2[0])}
    keys1, keys2 = sorted(map1), sorted(map2)
    k1, k2, F, epi = sc["k1"], sc["k2"], sc["F12s"], sc["epipoles"]
    m12 = [-1] * len(k1)
    hist = [[] for _ in range(30)]
    dist = lambda x, y: int(np.unpackbits(x ^ y).sum())

    def epipolar_ok(i1, i2, Fc):
        x1, y1, x2, y2 = f32(k1["x"][i1]), f32(k1["y"][i1]), f32(k2["x"][i2]), f32(k2["y"][i2])
        a = f32(f32(f32(x1 * Fc[0, 0]) + f32(y1 * Fc[1, 0])) + Fc[2, 0])
        b = f32(f32(f32(x1 * Fc[0, 1]) + f32(y1 * Fc[1, 1])) + Fc[2, 1])
        c = f32(f32(f32(x1 * Fc[0, 2]) + f32(y1 * Fc[1, 2])) + Fc[2, 2])
        num = f32(f32(f32(a * x2) + f32(b * y2)) + c)
        den = f32(f32(a * a) + f32(b * b))
        if den == 0:
            return False
        dsqr = f32(f32(num * num) / den)
        return float(dsqr) < 3.84 * float(sc["level_sigma2"][k2["octave"][i2]])

    nm, ia, ib = 0, 0, 0
    while ia < len(keys1) and ib < len(keys2):
        if keys1[ia] == keys2[ib]:
            for i1 in map1[keys1[ia]]:
                if sc["has_mp1"][i1] or not cam_enabled[sc["cam1"][i1]]:
                    continue
                st1 = sc["uright1"][i1] >= 0
                if only_stereo and not st1:
                    continue
                best, bi = 50, -1
                for i2 in map2[keys2[ib]]:
                    if sc["has_mp2"][i2] or sc["cam2"][i2] != sc["cam1"][i1]:
                        continue
                    st2 = sc["uright2"][i2] >= 0
                    if only_stereo and not st2:
                        continue
                    dd = dist(sc["d1"][i1], sc["d2"][i2])
                    if dd > 50 or dd > best:
                        continue
                    if not st1 and not st2:
                        c2 = sc["cam2"][i2]
                        ex = f32(epi[2 * c2] - f32(k2["x"][i2]))
                        ey = f32(epi[2 * c2 + 1] - f32(k2["y"][i2]))
                        if f32(f32(ex * ex) + f32(ey * ey)) < f32(f32(100) * sc["scale_factors"][k2["octave"][i2]]):
                            continue
                    if epipolar_ok(i1, i2, F[sc["cam1"][i1]]):
                        bi, best = i2, dd
                if bi >= 0:
                    m12[i1] = bi
                    nm += 1
                    if check_ori:
                        rot = f32(k1["angle"][i1]) - f32(k2["angle"][bi])
                        if rot < 0:
                            rot = f32(rot + f32(360.0))
                        b = int(math.floor(f32(rot * f32(1.0 / 30)) + 0.5))
                        hist[0 if b == 30 else b].append(i1)
            ia, ib = ia + 1, ib + 1
        elif keys1[ia] < keys2[ib]:
            ia = bisect.bisect_left(keys1, keys2[ib])
        else:
            ib = bisect.bisect_left(keys2, keys1[ia])
    if check_ori:
        import oracle_lib
        keep = set(oracle_lib.three_maxima([len(h) for h in hist]))
        for b in range(30):
            if b not in keep:
                for i1 in hist[b]:
                    m12[i1] = -1
                    nm -= 1
    return nm, m12


@pytest.mark.parametrize("seed,n_nodes,only_stereo,cam_enabled,check_ori", [
    (0, 10, False, (1, 1), True), (1, 30, False, (1, 0), True), (2, 6, True, (1, 1), False), (3, 15, False, (1, 1), False)])
def test_search_for_triangulation_vs_python(O, seed, n_nodes, only_stereo, cam_enabled, check_ori):
    from multi_orb_slam_b200.synth import feature_vector, triangulation_scene
    sc = triangulation_scene(160, 200, n_nodes, seed)
    node2 = np.where(sc["node2"] % 5 == 2, -1, sc["node2"])
    fv1, fv2 = feature_vector(sc["node1"]), feature_vector(node2)
    rn, rm12 = O.search_for_triangulation(sc, fv1, fv2, only_stereo, cam_enabled, check_ori)
    pn, pm12 = _py_search_for_triangulation(sc, fv1, fv2, only_stereo, cam_enabled, check_ori)
    assert rn == pn and list(rm12) == pm12
    assert rn == int((rm12 >= 0).sum())
    if not only_stereo:
        assert rn > 10


# ---- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:381-424) --------------------------
def _distinctive_sets(seed, sizes):
    from multi_orb_slam_b200.synth import perturbed_descriptors
    rng = np.random.default_rng(seed)
    blocks = []
    for n in sizes:
        base = random_descriptors(1, int(rng.integers(1 << 30)))
        if n:
            rows, _ = perturbed_descriptors(np.repeat(base, n, axis=0), int(rng.integers(1 << 30)), max_flips=60, permute=False)
            blocks.append(rows)
    desc = np.concatenate(blocks) if blocks else np.zeros((0, 32), np.uint8)
    return desc, np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)


def test_compute_distinctive_descriptors_vs_numpy(O):
    sizes = [1, 2, 3, 0, 7, 20, 33, 64, 5, 2, 100]
    desc, off = _distinctive_sets(0, sizes)
    got = O.compute_distinctive_descriptors(desc, off)
    for p, n in enumerate(sizes):
        if n == 0:
            assert got[p] == -1
            continue
        d = desc[off[p]:off[p + 1]]
        D = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(axis=2)
        med = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
        assert got[p] == int(np.argmin(med)), f"point {p} (N={n})"  # argmin = first minimum (:414-420)


# ---- DBoW2 vocabulary transform (TemplatedVocabulary.h:1127-1195, 1218-1259) ---------------------
def test_bow_transform_vs_python(O):
    from multi_orb_slam_b200.synth import random_vocabulary
    voc = random_vocabulary(6, 4, 3)
    rng = np.random.default_rng(4)
    leaves = np.nonzero(voc["word_id"] >= 0)[0]
    base = voc["node_desc"][rng.choice(leaves, 300)]
    bits = np.unpackbits(base, axis=1) ^ (rng.random((300, 256)) < 0.04).astype(np.uint8)
    desc = np.packbits(bits, axis=1)
    got = O.bow_transform(voc, desc, levelsup=2)
    cs, ci = voc["child_start"], voc["child_ids"]
    dist = lambda a, b: int(np.unpackbits(a ^ b).sum())
    bow, fv = {}, {}
    for i in range(len(desc)):
        node, level, nid = 0, 0, 0
        while cs[node] != cs[node + 1]:
            level += 1
            kids = ci[cs[node]:cs[node + 1]]
            best, node = dist(desc[i], voc["node_desc"][kids[0]]), int(kids[0])
            for c in kids[1:]:
                d = dist(desc[i], voc["node_desc"][c])
                if d < best:
                    best, node = d, int(c)
            if level == voc["L"] - 2:
                nid = node
        assert got["word"][i] == voc["word_id"][node] and got["node"][i] == nid and got["weight"][i] == voc["node_weight"][node]
        w = voc["node_weight"][node]
        if w > 0:
            bow[int(voc["word_id"][node])] = bow.get(int(voc["word_id"][node]), 0.0) + w
            fv.setdefault(nid, []).append(i)
    words = sorted(bow)
    norm = 0.0
    for wd in words:
        norm += abs(bow[wd])
    assert list(got["bow"][0]) == words and list(got["bow"][1]) == [bow[wd] / norm for wd in words]
    nodes = sorted(fv)
    assert list(got["featvec"][0]) == nodes
    for j, nd in enumerate(nodes):
        assert list(got["featvec"][2][got["featvec"][1][j]:got["featvec"][1][j + 1]]) == fv[nd]
    assert abs(sum(got["bow"][1]) - 1.0) < 1e-12 and len(nodes) > 5
