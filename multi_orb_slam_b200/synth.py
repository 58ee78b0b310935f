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
