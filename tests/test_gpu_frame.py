"""GPU parity tests of the Frame glue kernels (multi_orb_slam_b200/frame.py over the C-ABI) against the
oracle restatements of src/Frame.cc: UndistortKeyPoints (bit-exact with cv2 4.13 via the committed golden
vectors), ComputeImageBounds, ComputeStereoFromRGBD, AssignFeaturesToGrid — device-resident batches in the
extractor's output layout."""
import os

import numpy as np
import pytest

from multi_orb_slam_b200.synth import KP_DTYPE, camera_sequence

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TUM1 = (517.306408, 516.469215, 318.643040, 255.313989)
TUM1_DIST = (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)


def test_undistort_matches_cv2_golden():
    from multi_orb_slam_b200.frame import FrameGlue
    g = np.load(os.path.join(GOLD, "cv2_undistort.npz"))
    for i, c in enumerate(g["cams"]):
        glue = FrameGlue(c[0], c[1], c[2], c[3], c[4:9])
        k = np.zeros(len(g[f"pts_{i}"]), KP_DTYPE)
        k["x"], k["y"] = g[f"pts_{i}"][:, 0], g[f"pts_{i}"][:, 1]
        k["angle"], k["octave"] = 12.5, 3
        un = glue.UndistortKeyPoints(k)
        want = g[f"und_{i}"]
        assert np.array_equal(un["x"].view(np.uint32), want[:, 0].view(np.uint32)), f"camera {i}: x"
        assert np.array_equal(un["y"].view(np.uint32), want[:, 1].view(np.uint32)), f"camera {i}: y"
        assert (un["angle"] == 12.5).all() and (un["octave"] == 3).all()
        w, h = g[f"size_{i}"]
        b = glue.ComputeImageBounds(int(w), int(h))
        u = want[:4]
        assert (b.min_x, b.max_x, b.min_y, b.max_y) == (min(u[0, 0], u[2, 0]), max(u[1, 0], u[3, 0]), min(u[0, 1], u[1, 1]),
                                                        max(u[2, 1], u[3, 1]))


def test_no_distortion_is_identity(oracle_port):
    from multi_orb_slam_b200.frame import FrameGlue
    glue = FrameGlue(500, 500, 320, 240, (0, 0.2, 0.1, 0.1))  # k1 == 0 (:675-679, :772-778)
    k = np.zeros(10, KP_DTYPE)
    k["x"], k["y"] = np.arange(10) * 60.5, np.arange(10) * 40.25
    assert glue.UndistortKeyPoints(k).tobytes() == k.tobytes()
    b = glue.ComputeImageBounds(640, 480)
    assert (b.min_x, b.max_x, b.min_y, b.max_y) == (0.0, 640.0, 0.0, 480.0)


def test_extract_undistort_stereo_grid_on_device(oracle_port):
    """extractor -> UndistortKeyPoints -> ComputeStereoFromRGBD -> AssignFeaturesToGrid without leaving the GPU."""
    import torch
    from multi_orb_slam_b200.extractor import ORBextractor
    from multi_orb_slam_b200.frame import FrameGlue
    O = oracle_port
    F, W, H = 5, 640, 480
    imgs = camera_sequence(W, H, F, 77)
    ex = ORBextractor(1000, 1.2, 8, 20, 7, image_size=(W, H), max_batch=F)
    kps, desc, counts = ex.extract_batch_device(torch.from_numpy(imgs).cuda())
    ex.sync()
    glue = FrameGlue(*TUM1, TUM1_DIST, mbf=40.0)
    rng = np.random.default_rng(8)
    depth = np.where(rng.random((F, H, W)) < 0.75, rng.uniform(0.3, 9, (F, H, W)), 0).astype(np.float32)
    d_depth = torch.from_numpy(depth).cuda()
    kps_un = glue.undistort_batch_device(kps, counts)
    uright, dz = glue.stereo_from_rgbd_batch_device(kps, kps_un, counts, d_depth)
    bounds = glue.ComputeImageBounds(W, H)
    cell_start, items = glue.assign_features_to_grid_batch_device(kps_un, counts, bounds)
    glue.sync()
    n = counts.cpu().numpy()
    h_k, h_un = kps.cpu().numpy(), kps_un.cpu().numpy()
    h_ur, h_dz = uright.cpu().numpy(), dz.cpu().numpy()
    h_start, h_items = cell_start.cpu().numpy(), items.cpu().numpy().view(np.uint16)
    ob = O.compute_image_bounds(W, H, *TUM1, TUM1_DIST)
    assert (bounds.min_x, bounds.max_x, bounds.min_y, bounds.max_y) == ob
    for f in range(F):
        k = np.ascontiguousarray(h_k[f, : n[f]]).view(KP_DTYPE).reshape(-1)
        un = np.ascontiguousarray(h_un[f, : n[f]]).view(KP_DTYPE).reshape(-1)
        ref_un = O.undistort_keypoints(k, *TUM1, TUM1_DIST)
        assert un.tobytes() == ref_un.tobytes(), f"frame {f}: undistorted keypoints"
        r_ur, r_dz = O.compute_stereo_from_rgbd(k, ref_un, depth[f], 40.0)
        assert np.array_equal(h_ur[f, : n[f]], r_ur) and np.array_equal(h_dz[f, : n[f]], r_dz), f"frame {f}: stereo"
        assert (h_ur[f, n[f]:] == -1).all() and (h_dz[f, n[f]:] == -1).all()
        r_start, r_items = O.assign_features_to_grid(ref_un, ob)
        assert np.array_equal(h_start[f], r_start), f"frame {f}: grid cell starts"
        assert np.array_equal(h_items[f, : r_start[-1]].astype(np.int32), r_items), f"frame {f}: grid items"
    assert (h_dz > 0).sum() > 1000


@pytest.mark.parametrize("W,H,nf,F", [(640, 480, 1000, 3), (1241, 376, 2000, 2)])
def test_compute_stereo_matches_on_device(oracle_port, W, H, nf, F):
    """Frame::ComputeStereoMatches (src/Frame.cc:782-956): two extractors -> row-band Hamming association -> SAD
    refinement on the device-resident pyramids -> median cut, equal to the oracle bit for bit (mvuRight, mvDepth)."""
    import torch
    from multi_orb_slam_b200.extractor import ORBextractor
    from multi_orb_slam_b200.frame import FrameGlue
    from multi_orb_slam_b200.synth import stereo_pair
    O = oracle_port
    pairs = [stereo_pair(W, H, 40 + i, disparities=(3 + i, 11, 24 + 2 * i, 41)) for i in range(F)]
    exl = ORBextractor(nf, 1.2, 8, 20, 7, image_size=(W, H), max_batch=F)
    exr = ORBextractor(nf, 1.2, 8, 20, 7, image_size=(W, H), max_batch=F)
    dev_l = torch.from_numpy(np.stack([p[0] for p in pairs])).cuda()
    dev_r = torch.from_numpy(np.stack([p[1] for p in pairs])).cuda()
    kl, dl, nl = exl.extract_batch_device(dev_l)
    kr, dr, nr = exr.extract_batch_device(dev_r)
    glue = FrameGlue(718.856, 718.856, 607.1928, 185.2157, (0, 0, 0, 0, 0), mbf=386.1448)
    ur, z = glue.stereo_matches_batch_device(exl, exr, kl, dl, nl, kr, dr, nr)
    torch.cuda.synchronize()
    ur, z, nl = ur.cpu().numpy(), z.cpu().numpy(), nl.cpu().numpy()
    mb = np.float32(386.1448) / np.float32(718.856)
    n_ok = 0
    for f, (left, right) in enumerate(pairs):
        ref = []
        for img in (left, right):
            ex = O.extractor("port", nfeatures=nf)
            k, d, _ = ex.extract(img)
            ref.append((k, d, [ex.pyramid_level(l) for l in range(8)], ex.scale_tables()))
        (k0, d0, p0, t0), (k1, d1, p1, _) = ref
        assert nl[f] == len(k0)
        want_u, want_z = O.compute_stereo_matches(k0, d0, k1, d1, p0, p1, t0[0], t0[1], 386.1448, mb)
        assert np.array_equal(ur[f, : nl[f]].view(np.uint32), want_u.view(np.uint32)), f"pair {f}: mvuRight"
        assert np.array_equal(z[f, : nl[f]].view(np.uint32), want_z.view(np.uint32)), f"pair {f}: mvDepth"
        assert (ur[f, nl[f]:] == -1).all()
        n_ok += int((want_u >= 0).sum())
    assert n_ok > 0.3 * nl.sum()


def test_compute_stereo_matches_degenerate_pairs(oracle_port):
    """Empty / unmatched inputs: a constant right image (no right keypoints), a constant left image (no left
    keypoints), and a right image unrelated to the left one (few or no accepted matches: the median cut must not
    trip over an empty list) — all -1 or equal to the oracle."""
    import torch
    from multi_orb_slam_b200.extractor import ORBextractor
    from multi_orb_slam_b200.frame import FrameGlue
    from multi_orb_slam_b200.synth import textured
    O = oracle_port
    W, H, nf = 320, 240, 300
    a, b = textured(W, H, 71), textured(W, H, 72)
    flat = np.full((H, W), 90, np.uint8)
    lefts, rights = np.stack([a, flat, a, a]), np.stack([flat, a, b, a])
    exl = ORBextractor(nf, 1.2, 8, 20, 7, image_size=(W, H), max_batch=4)
    exr = ORBextractor(nf, 1.2, 8, 20, 7, image_size=(W, H), max_batch=4)
    dev_l, dev_r = torch.from_numpy(lefts).cuda(), torch.from_numpy(rights).cuda()  # level 0 of the pyramid views: keep alive
    kl, dl, nl = exl.extract_batch_device(dev_l)
    kr, dr, nr = exr.extract_batch_device(dev_r)
    glue = FrameGlue(400.0, 400.0, 160.0, 120.0, (0, 0, 0, 0, 0), mbf=40.0)
    ur, z = glue.stereo_matches_batch_device(exl, exr, kl, dl, nl, kr, dr, nr)
    torch.cuda.synchronize()
    ur, z, nl, nr = ur.cpu().numpy(), z.cpu().numpy(), nl.cpu().numpy(), nr.cpu().numpy()
    assert nr[0] == 0 and nl[1] == 0 and nl[0] > 100
    assert (ur[0] == -1).all() and (z[0] == -1).all() and (ur[1] == -1).all()
    mb = np.float32(40.0) / np.float32(400.0)
    for f in (2, 3):
        ref = []
        for img in (lefts[f], rights[f]):
            ex = O.extractor("port", nfeatures=nf)
            k, d, _ = ex.extract(img)
            ref.append((k, d, [ex.pyramid_level(l) for l in range(8)], ex.scale_tables()))
        (k0, d0, p0, t0), (k1, d1, p1, _) = ref
        want_u, want_z = O.compute_stereo_matches(k0, d0, k1, d1, p0, p1, t0[0], t0[1], 40.0, mb)
        assert np.array_equal(ur[f, : nl[f]].view(np.uint32), want_u.view(np.uint32)), f"pair {f}"
        assert np.array_equal(z[f, : nl[f]].view(np.uint32), want_z.view(np.uint32)), f"pair {f}"
    # identical images: every SAD is 0, so the median threshold is 0 and the cut of :945-955 removes every match
    assert (ur[3, : nl[3]] == -1).all()
