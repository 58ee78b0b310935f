"""GPU parity tests: the CUDA extractor (through the C-ABI) against the CPU oracle on identical
seeded frames.  Bit-exact for pyramid bytes, FAST candidates, blurred images, keypoint (x, y,
octave, size, response) and descriptors; angle within 1e-4 relative (BASELINE.json north_star)
— in practice it is bit-identical too, which the test reports."""
import numpy as np
import pytest

from multi_orb_slam_b200.synth import textured

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def O(oracle_port):
    return oracle_port


def _gpu(nfeatures=1000, size=(640, 480), max_batch=1, **kw):
    from multi_orb_slam_b200.extractor import ORBextractor
    return ORBextractor(nfeatures, kw.get("scale", 1.2), kw.get("nlevels", 8), kw.get("ini", 20), kw.get("min", 7),
                        image_size=size, max_batch=max_batch)


def _assert_same_features(k_gpu, d_gpu, k_ref, d_ref, tag=""):
    assert len(k_gpu) == len(k_ref), f"{tag}: count {len(k_gpu)} vs {len(k_ref)}"
    for f in ("x", "y", "size", "response", "octave"):
        assert np.array_equal(k_gpu[f], k_ref[f]), f"{tag}: field {f} differs"
    np.testing.assert_allclose(k_gpu["angle"], k_ref["angle"], rtol=1e-4, atol=0, err_msg=f"{tag}: angle")
    assert np.array_equal(d_gpu, d_ref), f"{tag}: {(d_gpu != d_ref).any(axis=1).sum()} descriptor rows differ"


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_stages_640x480(O, seed):
    img = textured(640, 480, seed)
    port = O.extractor("port")
    k_ref, d_ref, c_ref = port.extract(img)
    ex = _gpu()
    k, d = ex(img)
    for l in range(8):
        assert np.array_equal(ex.pyramid_level(l, with_border=True), port.pyramid_level(l)), f"pyramid level {l}"
    for l in range(8):
        gx, gy, gs = ex.debug_candidates(l)
        rx, ry, rs = port.candidates(l)
        assert len(gx) == len(rx), f"level {l}: {len(gx)} candidates vs {len(rx)}"
        assert np.array_equal(gx, rx) and np.array_equal(gy, ry) and np.array_equal(gs, rs), f"candidates level {l}"
    for l in range(8):
        lv = ex.pyramid_level(l)
        ref_blur = port.blurred(l, lv.shape[1], lv.shape[0])
        if ref_blur is not None:
            assert np.array_equal(ex.debug_blurred(l), ref_blur), f"blur level {l}"
    _assert_same_features(k, d, k_ref, d_ref, f"seed {seed}")
    assert np.array_equal(k["angle"], k_ref["angle"]), "angles are expected to be bit-identical"


@pytest.mark.parametrize("size,nfeatures", [((1241, 376), 2000), ((1280, 720), 1000), ((640, 480), 500),
                                            ((752, 480), 1200), ((320, 240), 300)])
def test_other_geometries(O, size, nfeatures):
    img = textured(size[0], size[1], 11)
    k_ref, d_ref, _ = O.extractor("port", nfeatures=nfeatures).extract(img)
    k, d = _gpu(nfeatures, size)(img)
    _assert_same_features(k, d, k_ref, d_ref, f"{size} nf={nfeatures}")


def test_low_texture_uses_min_threshold(O):
    """Flat image with faint texture: most cells fall back to minThFAST (ORBextractor.cc:813-817)."""
    rng = np.random.default_rng(3)
    img = (128 + rng.integers(-6, 7, size=(480, 640))).astype(np.uint8)
    img[200:280, 300:380] += 40
    k_ref, d_ref, _ = O.extractor("port").extract(img)
    k, d = _gpu()(img)
    _assert_same_features(k, d, k_ref, d_ref, "low texture")


def test_keypoints_only_in_the_right_half(O):
    """Regression (found by tools/dev/fuzz_parity.py under compute-sanitizer): when the LEFT octree roots of a level hold no
    key, lanes past the last key of k_octree's key pass followed label 0 into a move-table entry nobody had written and
    indexed the node array with whatever shared memory held.  Texture in the right third only, on a wide frame (four roots)."""
    w, h = 1241, 376
    rng = np.random.default_rng(17)
    img = (128 + rng.integers(-5, 6, size=(h, w))).astype(np.uint8)
    img[:, 2 * w // 3:] = textured(w, h, 4)[:, 2 * w // 3:]
    ref = O.extractor("port", nfeatures=1733)
    k_ref, d_ref, _ = ref.extract(img)
    ex = _gpu(1733, (w, h), max_batch=3)
    for _ in range(3):  # shared memory of earlier launches is what the stray index used to come from
        k, d = ex(img)
        _assert_same_features(k, d, k_ref, d_ref, "right half only")
    assert len(k_ref) > 100


def test_constant_image_gives_no_keypoints():
    k, d = _gpu()(np.full((480, 640), 77, dtype=np.uint8))
    assert len(k) == 0 and d.shape == (0, 32)


def test_empty_image_returns_silently():
    k, d = _gpu()(np.zeros((0, 0), dtype=np.uint8))
    assert len(k) == 0


def test_strided_input(O):
    big = textured(800, 480, 5)
    view = big[:, 80:720]  # row stride 800
    k_ref, d_ref, _ = O.extractor("port").extract(np.ascontiguousarray(view))
    k, d = _gpu()(view)
    _assert_same_features(k, d, k_ref, d_ref, "strided")


def test_batch_matches_single_and_oracle(O):
    F = 12
    imgs = np.stack([textured(640, 480, 100 + i) for i in range(F)])
    ex = _gpu(max_batch=5)  # 12 frames through a 5-frame workspace: 3 launch groups
    kps, desc, counts = ex.extract_batch(imgs)
    port = O.extractor("port")
    for f in range(F):
        k_ref, d_ref, _ = port.extract(imgs[f])
        _assert_same_features(kps[f, : counts[f]], desc[f, : counts[f]], k_ref, d_ref, f"frame {f}")


def test_device_batch_api(O):
    import torch
    F = 4
    imgs = np.stack([textured(640, 480, 200 + i) for i in range(F)])
    ex = _gpu(max_batch=4)
    t = torch.from_numpy(imgs).cuda()
    torch.cuda.synchronize()
    kps, desc, counts = ex.extract_batch_device(t)
    ex.sync()
    kps, desc, counts = kps.cpu().numpy(), desc.cpu().numpy(), counts.cpu().numpy()
    from multi_orb_slam_b200._lib import KP_DTYPE
    port = O.extractor("port")
    for f in range(F):
        k = kps[f, : counts[f]].copy().view(KP_DTYPE).reshape(-1)
        k_ref, d_ref, _ = port.extract(imgs[f])
        _assert_same_features(k, desc[f, : counts[f]], k_ref, d_ref, f"device frame {f}")


def test_scale_tables_and_quotas(O):
    ex = _gpu()
    port = O.extractor("port")
    for a, b in zip(ex._tables(), port.scale_tables()):
        assert np.array_equal(a, b)
    assert np.array_equal(ex.features_per_level(), port.features_per_level())
    assert list(ex.features_per_level()) == [217, 181, 151, 126, 105, 87, 73, 60]


def test_capacity_error():
    from multi_orb_slam_b200 import _lib
    import ctypes as C
    ex = _gpu()
    img = textured(640, 480, 0)
    kps = np.empty(10, dtype=_lib.KP_DTYPE)
    desc = np.empty((10, 32), dtype=np.uint8)
    n = C.c_int()
    rc = _lib.lib.orbx_extract(ex._h, img.ctypes.data, 480, 640, 640, kps.ctypes.data, desc.ctypes.data, 10, C.byref(n))
    assert rc == _lib.E_CAPACITY and n.value == 10


def test_idempotent_and_deterministic():
    img = textured(640, 480, 42)
    ex = _gpu()
    k1, d1 = ex(img)
    k2, d2 = ex(img)
    assert np.array_equal(k1, k2) and np.array_equal(d1, d2)


def test_two_handles_with_different_nfeatures(O):
    """Regression: the octree kernel's shared-memory opt-in is per function, not per handle."""
    img = textured(640, 480, 8)
    big, small = _gpu(1000), _gpu(500)
    for ex, nf in ((big, 1000), (small, 500), (big, 1000)):
        k_ref, d_ref, _ = O.extractor("port", nfeatures=nf).extract(img)
        k, d = ex(img)
        _assert_same_features(k, d, k_ref, d_ref, f"nf={nf}")


@pytest.mark.parametrize("mode", ["bands", "split"])
@pytest.mark.parametrize("case", ["textured", "kitti", "hd", "small", "low_texture", "noise", "constant"])
def test_band_fast_kernel_matches_the_oracle(O, monkeypatch, case, mode):
    """The other two forms of the FAST stage, chosen per handle with ORB_B200_FAST: `bands` (k_fast_bands, the whole stage
    streamed per band) and `split` (k_fast_prefilter writes the rejection test's pass bits, k_fast_cells starts at the arc
    measure).  Candidate lists per level incl. their order, and the final features, equal to the oracle
    — on textured frames of several geometries (whole and clipped cells, 3..8 cells per band), on a low-texture frame
    (most cells rerun at minThFAST), on uniform noise (dense candidates: the bounded buffers fill) and on a constant
    frame (cells that keep nothing in either pass)."""
    monkeypatch.setenv("ORB_B200_FAST", mode)
    size, nf = {"kitti": ((1241, 376), 2000), "hd": ((1280, 720), 1000), "small": ((320, 240), 300)}.get(case, ((640, 480), 1000))
    w, h = size
    if case == "low_texture":
        rng = np.random.default_rng(3)
        img = (128 + rng.integers(-6, 7, size=(h, w))).astype(np.uint8)
        img[200:280, 300:380] += 40
    elif case == "noise":
        img = np.random.default_rng(5).integers(0, 256, size=(h, w), dtype=np.uint8)
    elif case == "constant":
        img = np.full((h, w), 77, dtype=np.uint8)
    else:
        img = textured(w, h, 21)
    port = O.extractor("port", nfeatures=nf)
    k_ref, d_ref, _ = port.extract(img)
    ex = _gpu(nf, size)
    k, d = ex(img)
    for l in range(8):
        gx, gy, gs = ex.debug_candidates(l)
        rx, ry, rs = port.candidates(l)
        assert len(gx) == len(rx), f"level {l}: {len(gx)} candidates vs {len(rx)}"
        assert np.array_equal(gx, rx) and np.array_equal(gy, ry) and np.array_equal(gs, rs), f"candidates level {l}"
    _assert_same_features(k, d, k_ref, d_ref, case)


def test_full_size_batch_is_frame_independent(O):
    """BASELINE.json configs[1] size — 256 frames per handle, both camera quotas — through the size-independent
    property of the path: camera-frames are independent units, so a batch of 8 distinct frames repeated in scrambled
    order must return, in EVERY slot, the oracle's features for the frame in that slot, and consecutive-frame
    SearchForInitialization over the batch must equal the oracle's result for the same pair of frames."""
    import torch
    from multi_orb_slam_b200._lib import KP_DTYPE, Bounds
    from multi_orb_slam_b200.matcher import ORBmatcher
    F, D = 256, 8
    base = np.stack([textured(640, 480, 300 + i) for i in range(D)])
    order = np.random.default_rng(12).integers(0, D, F)
    dev = torch.from_numpy(base[order]).cuda()
    for nf in (1000, 500):
        ex = _gpu(nfeatures=nf, max_batch=F)
        kps, desc, counts = ex.extract_batch_device(dev)
        ex.sync()
        h_k, h_d, h_n = kps.cpu().numpy(), desc.cpu().numpy(), counts.cpu().numpy()
        port = O.extractor("port", nfeatures=nf)
        ref = [port.extract(base[i])[:2] for i in range(D)]
        for f in range(F):
            k = h_k[f, : h_n[f]].copy().view(KP_DTYPE).reshape(-1)
            _assert_same_features(k, h_d[f, : h_n[f]], *ref[order[f]], f"nfeatures {nf} slot {f} (frame {order[f]})")
        if nf != 1000:
            continue
        m = ORBmatcher(0.9, True)
        cap = kps.shape[1]
        m12 = torch.empty((F - 1, cap), dtype=torch.int32, device="cuda")
        nm = torch.empty((F - 1,), dtype=torch.int32, device="cuda")
        m.search_for_initialization_device(F - 1, cap, kps, desc, counts, kps[1:], desc[1:], counts[1:],
                                           Bounds(0.0, 640.0, 0.0, 480.0), None, 100, m12, nm)
        torch.cuda.synchronize()
        m12, nm = m12.cpu().numpy(), nm.cpu().numpy()
        cache = {}
        for f in range(F - 1):
            key = (int(order[f]), int(order[f + 1]))
            if key not in cache:
                (k1, d1), (k2, d2) = ref[key[0]], ref[key[1]]
                prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
                cache[key] = O.search_for_initialization(k1, d1, k2, d2, (0, 640, 0, 480), prev, 100, 0.9, True)
            rn, rm12, _ = cache[key]
            assert nm[f] == rn and np.array_equal(m12[f, : len(rm12)], rm12), f"pair {f} {key}"
