"""Seeded synthetic inputs shared by the tests, the bench and the CPU baseline.

Everything is generated on the CPU with numpy only, so the GPU path and the oracle always see
identical bytes.  No reference data set ships with AlterPang/Multi_ORB_SLAM (README.md:52-73
asks users to record their own), hence synthetic frames.
"""
from __future__ import annotations

import numpy as np


def _cubic_upsample(a: np.ndarray, f: int) -> np.ndarray:
    """Catmull-Rom (a=-0.5) upsampling by integer factor f along both axes, edge-replicated."""
    if f == 1:
        return a
    t = (np.arange(f, dtype=np.float64) + 0.5) / f - 0.5  # sample offsets within a source cell
    def weights(t):
        t = np.abs(t)
        return np.where(t <= 1, 1.5 * t**3 - 2.5 * t**2 + 1,
                        np.where(t < 2, -0.5 * t**3 + 2.5 * t**2 - 4 * t + 2, 0.0))
    def up1(x):  # along axis 0
        n = x.shape[0]
        out = np.zeros((n * f,) + x.shape[1:], dtype=np.float64)
        for j in range(f):
            base = np.floor(t[j]).astype(int)
            frac = t[j] - base
            acc = 0
            for k in range(-1, 3):
                idx = np.clip(np.arange(n) + base + k, 0, n - 1)
                acc = acc + weights(frac - k) * x[idx]
            out[j::f] = acc
        return out
    return up1(up1(a).T).T


def textured(width: int, height: int, seed: int) -> np.ndarray:
    """Textured u8 frame: 5 octaves of cubic-upsampled uniform noise (cells 1,2,4,8,16 px,
    amplitudes 4,20,50,60,40) + 150 random rectangles (side 6-60 px, offset U[-70,70]),
    min-max normalised.  The fine-octave amplitudes are lower than SURVEY.md §8d's first
    guess so that level 0 yields ~4-5 k FAST candidates at 640x480 (camera-like density)
    instead of ~12 k."""
    rng = np.random.default_rng(seed)
    img = np.zeros((height, width), dtype=np.float64)
    for cell, amp in zip((1, 2, 4, 8, 16), (4.0, 20.0, 50.0, 60.0, 40.0)):
        h, w = -(-height // cell), -(-width // cell)
        noise = rng.uniform(-1.0, 1.0, size=(h, w))
        img += amp * _cubic_upsample(noise, cell)[:height, :width]
    for _ in range(150):
        rw, rh = rng.integers(6, 61, size=2)
        x0 = int(rng.integers(0, max(1, width - rw)))
        y0 = int(rng.integers(0, max(1, height - rh)))
        img[y0:y0 + rh, x0:x0 + rw] += rng.uniform(-70.0, 70.0)
    lo, hi = img.min(), img.max()
    return np.clip(np.rint((img - lo) * (255.0 / (hi - lo))), 0, 255).astype(np.uint8)


def shifted_noisy(img: np.ndarray, seed: int, max_shift: int = 8, sigma: float = 2.0) -> np.ndarray:
    """Next frame of a synthetic sequence (config 2): integer shift in [-max_shift, max_shift]^2
    (edge-replicated) plus N(0, sigma) noise."""
    rng = np.random.default_rng(seed)
    dx, dy = (int(v) for v in rng.integers(-max_shift, max_shift + 1, size=2))
    h, w = img.shape
    ys = np.clip(np.arange(h) - dy, 0, h - 1)
    xs = np.clip(np.arange(w) - dx, 0, w - 1)
    out = img[np.ix_(ys, xs)].astype(np.float64) + rng.normal(0.0, sigma, size=img.shape)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def camera_sequence(width: int, height: int, n_frames: int, seed: int, max_shift: int = 8, sigma: float = 2.0):
    """n_frames of one synthetic camera stream (config 2): the clean scene moves by an integer
    offset in [-max_shift, max_shift]^2 per frame (wrapping around, so the texture statistics
    stay stationary over long sequences) and every frame carries FRESH N(0, sigma) sensor noise
    (noise does not accumulate from frame to frame)."""
    rng = np.random.default_rng(100003 * seed + 17)
    clean = textured(width, height, seed)
    frames = []
    for _ in range(n_frames):
        noisy = clean.astype(np.float32) + rng.normal(0.0, sigma, size=clean.shape).astype(np.float32)
        frames.append(np.clip(np.rint(noisy), 0, 255).astype(np.uint8))
        dx, dy = (int(v) for v in rng.integers(-max_shift, max_shift + 1, size=2))
        clean = np.roll(clean, (dy, dx), axis=(0, 1))
    return np.stack(frames)


def stereo_pair(width: int, height: int, seed: int, disparities=(4, 9, 17, 30), sigma: float = 1.5):
    """Rectified stereo pair: the right image is the left one with horizontal bands shifted by the band's
    disparity (a feature at column u of the left image sits at u - d in the right one) plus N(0, sigma) noise."""
    left = textured(width + 64, height, seed)
    right = np.empty((height, width), np.uint8)
    rng = np.random.default_rng(seed + 500)
    edges = np.linspace(0, height, len(disparities) + 1).astype(int)
    for d, y0, y1 in zip(disparities, edges[:-1], edges[1:]):
        right[y0:y1] = left[y0:y1, d:d + width]
    right = np.clip(np.rint(right.astype(np.float32) + rng.normal(0, sigma, right.shape)), 0, 255).astype(np.uint8)
    return np.ascontiguousarray(left[:, :width]), right


def random_descriptors(n: int, seed: int) -> np.ndarray:
    """n x 32 bytes of i.i.d. bits (config 3, set A)."""
    return np.random.default_rng(seed).integers(0, 256, size=(n, 32), dtype=np.uint8)


def perturbed_descriptors(desc: np.ndarray, seed: int, max_flips: int = 80, permute: bool = True):
    """Set B of config 3: rows of `desc` (optionally permuted) with k~U{0..max_flips} random bit
    flips each.  Returns (B, perm) with B[i] derived from desc[perm[i]]."""
    rng = np.random.default_rng(seed)
    n = desc.shape[0]
    perm = rng.permutation(n) if permute else np.arange(n)
    bits = np.unpackbits(desc[perm], axis=1)
    k = rng.integers(0, max_flips + 1, size=n)
    # choose flip positions: rank of uniform noise < k  (k distinct positions per row)
    r = rng.random((n, 256))
    flip = np.argsort(np.argsort(r, axis=1), axis=1) < k[:, None]
    bits ^= flip.astype(np.uint8)
    return np.packbits(bits, axis=1), perm


def feature_vector(node_of_feature: np.ndarray):
    """DBoW2::FeatureVector (std::map<node id, vector<feature index>>) flattened to CSR:
    (node_ids ascending, start[nn + 1], items) with the items of a node in ascending feature
    index, the order DBoW2's transform() appends them.  Features with node id < 0 are left out."""
    node_of_feature = np.asarray(node_of_feature, dtype=np.int64)
    idx = np.nonzero(node_of_feature >= 0)[0]
    order = idx[np.argsort(node_of_feature[idx], kind="stable")]
    nodes, counts = np.unique(node_of_feature[order], return_counts=True)
    start = np.zeros(len(nodes) + 1, dtype=np.int32)
    start[1:] = np.cumsum(counts)
    return nodes.astype(np.int32), start, order.astype(np.int32)


def bow_scene(n1: int, n2: int, n_nodes: int, seed: int, p_same_node: float = 0.85, max_flips: int = 70):
    """Two feature sets for SearchByBoW: side 2 holds noisy copies of random side-1 features, most
    of them in the same vocabulary node.  Returns dict(d1, a1, node1, d2, a2, node2, src)."""
    rng = np.random.default_rng(seed)
    d1 = random_descriptors(n1, seed + 1)
    a1 = rng.uniform(0, 360, n1).astype(np.float32)
    node1 = rng.integers(0, n_nodes, n1)
    src = rng.integers(0, n1, n2)
    bits = np.unpackbits(d1[src], axis=1)
    flips = rng.integers(0, max_flips + 1, n2)
    bits ^= (np.argsort(np.argsort(rng.random((n2, 256)), axis=1), axis=1) < flips[:, None]).astype(np.uint8)
    d2 = np.packbits(bits, axis=1)
    a2 = ((a1[src] + rng.normal(0, 15, n2)) % 360).astype(np.float32)
    node2 = np.where(rng.random(n2) < p_same_node, node1[src], rng.integers(0, n_nodes, n2))
    return dict(d1=d1, a1=a1, node1=node1, d2=d2, a2=a2, node2=node2, src=src)


# orbx_keypoint rows (include/orb_b200.h), same as _lib.KP_DTYPE (not imported: synth must work without the library)
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])


def triangulation_scene(n1: int, n2: int, n_nodes: int, seed: int, n_levels: int = 8):
    """Two two-camera key frames observing the same 3-D points from slightly different poses, for
    SearchForTriangulation: side 2 holds noisy re-observations of random side-1 features (same
    camera, mostly the same vocabulary node).  Returns a dict of the flat arrays the C-ABI takes,
    including the two fundamental matrices and epipoles computed the way the reference does
    (F12 = K^-T [t12]x R12 K^-1, src/ORBmatcher.cc:1421-1423) in float64, rounded to float32."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = 520.0, 520.0, 320.0, 240.0
    K = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    Kinv = np.linalg.inv(K)

    def skew(t):
        return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])

    def rot(rx, ry, rz):
        cxr, sxr, cyr, syr, czr, szr = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
        Rx = np.array([[1, 0, 0], [0, cxr, -sxr], [0, sxr, cxr]])
        Ry = np.array([[cyr, 0, syr], [0, 1, 0], [-syr, 0, cyr]])
        Rz = np.array([[czr, -szr, 0], [szr, czr, 0], [0, 0, 1]])
        return Rz @ Ry @ Rx

    # per camera: X1 = R12 X2 + t12 (coordinates of a point in key frame 1 from key frame 2)
    R12 = [rot(0.01, -0.02, 0.015), rot(-0.015, 0.01, 0.02)]
    t12 = [np.array([0.25, 0.02, 0.03]), np.array([0.03, 0.02, 0.28])]
    F12s = np.stack([Kinv.T @ skew(t12[c]) @ R12[c] @ Kinv for c in range(2)])
    epi = []
    for c in range(2):
        C2 = -R12[c].T @ t12[c]  # centre of camera 1 in the frame of camera 2
        epi += [fx * C2[0] / C2[2] + cx, fy * C2[1] / C2[2] + cy]
    cam1 = rng.integers(0, 2, n1)
    X2 = np.stack([rng.uniform(-2, 2, n1), rng.uniform(-1.5, 1.5, n1), rng.uniform(2, 8, n1)], axis=1)
    X1 = np.stack([R12[c] @ x + t12[c] for c, x in zip(cam1, X2)])
    p1 = (K @ (X1 / X1[:, 2:3]).T).T[:, :2]
    p2_true = (K @ (X2 / X2[:, 2:3]).T).T[:, :2]
    k1 = np.zeros(n1, KP_DTYPE)
    k1["x"], k1["y"] = p1[:, 0], p1[:, 1]
    k1["octave"] = rng.integers(0, 4, n1)
    k1["angle"] = rng.uniform(0, 360, n1)
    d1 = random_descriptors(n1, seed + 1)
    node1 = rng.integers(0, n_nodes, n1)
    src = rng.integers(0, n1, n2)
    k2 = np.zeros(n2, KP_DTYPE)
    noise = rng.normal(0, 1.0, (n2, 2)) * np.where(rng.random(n2) < 0.8, 0.4, 6.0)[:, None]
    k2["x"], k2["y"] = p2_true[src, 0] + noise[:, 0], p2_true[src, 1] + noise[:, 1]
    k2["octave"] = np.clip(k1["octave"][src] + rng.integers(-1, 2, n2), 0, n_levels - 1)
    k2["angle"] = (k1["angle"][src] + rng.normal(0, 12, n2)) % 360
    bits = np.unpackbits(d1[src], axis=1)
    flips = rng.integers(0, 60, n2)
    bits ^= (np.argsort(np.argsort(rng.random((n2, 256)), axis=1), axis=1) < flips[:, None]).astype(np.uint8)
    d2 = np.packbits(bits, axis=1)
    cam2 = np.where(rng.random(n2) < 0.9, cam1[src], 1 - cam1[src])
    node2 = np.where(rng.random(n2) < 0.85, node1[src], rng.integers(0, n_nodes, n2))
    sf = (1.2 ** np.arange(n_levels)).astype(np.float32)
    return dict(k1=k1, d1=d1, has_mp1=(rng.random(n1) < 0.3).astype(np.int32), cam1=cam1.astype(np.int32),
                uright1=np.where(rng.random(n1) < 0.2, k1["x"] - 5, -1).astype(np.float32), node1=node1,
                k2=k2, d2=d2, has_mp2=(rng.random(n2) < 0.3).astype(np.int32), cam2=cam2.astype(np.int32),
                uright2=np.where(rng.random(n2) < 0.2, k2["x"] - 5, -1).astype(np.float32), node2=node2,
                F12s=F12s.astype(np.float32), epipoles=np.array(epi, dtype=np.float32), scale_factors=sf,
                level_sigma2=(sf * sf).astype(np.float32), src=src)


def random_vocabulary(k: int, L: int, seed: int, p_stop: float = 0.02, p_short: float = 0.05):
    """A DBoW2-shaped vocabulary tree for tests (the real ORBvoc.txt, k = 10, L = 6, is not shipped with the
    reference): node 0 is the root, every inner node has 2..k children whose descriptors are noisy copies of the
    parent's (like k-means centres of a cluster), a few branches end early (leaves above depth L, as DBoW2
    produces for small clusters), a few words carry weight 0 (stopped).  Nodes are numbered in creation order,
    depth first like TemplatedVocabulary::HKmeansStep; words in leaf creation order."""
    rng = np.random.default_rng(seed)
    desc, children, depth = [rng.integers(0, 256, 32, dtype=np.uint8)], [[]], [0]

    def grow(node):
        if depth[node] == L or (depth[node] >= 1 and rng.random() < p_short):
            return
        n_child = k if rng.random() < 0.8 else int(rng.integers(2, k + 1))
        ids = []
        for _ in range(n_child):
            bits = np.unpackbits(desc[node])
            flip = rng.random(256) < 0.5 ** (depth[node] + 1) * 0.6
            desc.append(np.packbits(bits ^ flip.astype(np.uint8)))
            children.append([])
            depth.append(depth[node] + 1)
            ids.append(len(desc) - 1)
        children[node] = ids
        for c in ids:
            grow(c)

    import sys
    sys.setrecursionlimit(10000)
    grow(0)
    n = len(desc)
    child_start = np.zeros(n + 1, dtype=np.int32)
    child_start[1:] = np.cumsum([len(c) for c in children])
    child_ids = np.array([c for cs in children for c in cs], dtype=np.int32)
    word_id = np.full(n, -1, dtype=np.int32)
    leaves = [i for i in range(n) if not children[i]]
    word_id[leaves] = np.arange(len(leaves))
    weight = np.zeros(n, dtype=np.float64)
    weight[leaves] = np.where(rng.random(len(leaves)) < p_stop, 0.0, rng.uniform(0.5, 9.0, len(leaves)))
    return dict(child_start=child_start, child_ids=child_ids, node_desc=np.stack(desc), word_id=word_id, node_weight=weight,
                L=L, depth=np.array(depth))


# ---- two-camera rig scene for the pose-based searches (SearchByProjection overloads, Fuse, SearchBySim3) ----
RIG_CALIB = np.array([[0.01001086, 0.01371197, 0.99975906], [0.02114039, 0.99964902, -0.01393279],
                  [-0.99963218, 0.02128624, 0.00971428], [0.1609449, 0.00377988, -0.07087293]], dtype=np.float32)  # OtherFiles/calibration.txt
RIG_CAM = (517.3, 516.5, 318.6, 255.3, 0.08, 40.0)  # fx fy cx cy mb mbf


def rig_rotation(rx, ry, rz):
    cx, sx, cy, sy, cz, sz = np.cos(rx), np.sin(rx), np.cos(ry), np.sin(ry), np.cos(rz), np.sin(rz)
    return (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @
            np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))


def rig_scene(extract, seed, n_last, tlast_offset):
    """Two-camera current frame + last-frame map points that project near current keypoints.
    extract(nfeatures, image) -> (keypoints[KP_DTYPE], descriptors) supplies the features (the oracle in the
    tests, the CUDA extractor in bench.py)."""
    rng = np.random.default_rng(seed)
    k0, d0 = extract(1000, textured(640, 480, seed))
    k1, d1 = extract(500, textured(640, 480, seed + 100))
    cur_k, cur_d = np.concatenate([k0, k1]), np.concatenate([d0, d1])
    cur_cam = np.concatenate([np.zeros(len(k0), np.int32), np.ones(len(k1), np.int32)])
    n = len(cur_k)
    fx, fy, cx, cy, mb, mbf = RIG_CAM
    Tcw = np.eye(4)
    Tcw[:3, :3] = rig_rotation(0.02, -0.03, 0.01)
    Tcw[:3, 3] = [0.05, -0.02, 0.1]
    Tlw = np.eye(4)
    Tlw[:3, :3] = rig_rotation(0.0, 0.0, 0.0)
    Tlw[:3, 3] = np.array([0.05, -0.02, 0.1]) + np.array(tlast_offset)
    R12, t12 = RIG_CALIB[:3].astype(np.float64), RIG_CALIB[3].astype(np.float64)
    R21, t21 = R12.T, -R12.T @ t12
    src = rng.integers(0, n, n_last)
    z = rng.uniform(1.0, 8.0, n_last)
    u = cur_k["x"][src] + rng.normal(0, 2.5, n_last)
    v = cur_k["y"][src] + rng.normal(0, 2.5, n_last)
    Xc = np.stack([(u - cx) / fx * z, (v - cy) / fy * z, z], axis=1)  # in the source keypoint's camera frame
    is1 = cur_cam[src] == 1
    Xc0 = np.where(is1[:, None], (Xc - t21) @ R21, Xc)               # back to the rig (camera 0) frame: R21^T (x - t21)
    Xw = (Xc0 - Tcw[:3, 3]) @ Tcw[:3, :3]                            # R^T (x - t)
    Xw[rng.random(n_last) < 0.05] *= -1                               # some behind the camera
    bits = np.unpackbits(cur_d[src], axis=1)
    flips = rng.integers(0, 70, n_last)
    bits ^= (np.argsort(np.argsort(rng.random((n_last, 256)), axis=1), axis=1) < flips[:, None]).astype(np.uint8)
    last_k = np.zeros(n_last, KP_DTYPE)
    last_k["octave"] = np.clip(cur_k["octave"][src] + rng.integers(-1, 2, n_last), 0, 7)
    last_k["angle"] = (cur_k["angle"][src] + rng.normal(0, 15, n_last)) % 360
    ur = np.where(rng.random(n) < 0.7, cur_k["x"] - mbf / rng.uniform(1, 8, n), -1).astype(np.float32)
    return dict(cur_k=cur_k, cur_d=cur_d, cur_cam=cur_cam, ur=ur, Tcw=Tcw.astype(np.float32), Tlw=Tlw.astype(np.float32),
                last_k=last_k, last_cam=cur_cam[src].copy(), last_valid=(rng.random(n_last) < 0.9).astype(np.int32),
                last_xyz=Xw.astype(np.float32), last_desc=np.packbits(bits, axis=1),
                last_obs=(rng.random(n_last) < 0.85).astype(np.int32), rng=rng, n=n)


def projection_case(k, d, seed, nmp, width=1241, height=376):
    """BASELINE configs[3]'s matcher input: `nmp` map points for SearchByProjection(Frame, vector<MapPoint*>), derived
    from a frame's keypoints `k` / descriptors `d` (most project near the keypoint they come from, with up to 60 flipped
    descriptor bits; some project anywhere).  Returns (mp records, mp descriptors, rng)."""
    from ._lib import MP_DTYPE
    rng = np.random.default_rng(seed + 50)
    mp = np.zeros(nmp, MP_DTYPE)
    src = rng.integers(0, len(k), nmp)
    near = rng.random(nmp) < 0.8
    mp["proj_x"] = np.where(near, k["x"][src] + rng.normal(0, 3, nmp), rng.uniform(0, width, nmp)).astype(np.float32)
    mp["proj_y"] = np.where(near, k["y"][src] + rng.normal(0, 3, nmp), rng.uniform(0, height, nmp)).astype(np.float32)
    mp["proj_xr"] = mp["proj_x"] - 5
    mp["view_cos"] = rng.uniform(0.5, 1.0, nmp).astype(np.float32)
    mp["view_cos"][rng.random(nmp) < 0.1] = 0.9995
    mp["level"] = np.where(near, np.clip(k["octave"][src] + rng.integers(0, 2, nmp), 0, 7), rng.integers(0, 8, nmp))
    mp["track_in_view"] = rng.random(nmp) < 0.95
    mp["bad"] = rng.random(nmp) < 0.03
    bits = np.unpackbits(d[src], axis=1)
    flips = rng.integers(0, 61, nmp)
    bits ^= (np.argsort(np.argsort(rng.random((nmp, 256)), axis=1), axis=1) < flips[:, None]).astype(np.uint8)
    return mp, np.packbits(bits, axis=1), rng
