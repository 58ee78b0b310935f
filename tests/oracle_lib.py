"""ctypes bindings for the CPU oracle (TEST INFRASTRUCTURE).

Loaded only by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs — never by the
product package.  `port` = oracle/liborb_oracle.so (restatement, built from oracle/*.cc);
`ref` = oracle/_ref/liborb_ref.so (reference ORBextractor.cc compiled verbatim; prebuilt in the
authoring container, may be absent elsewhere)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])
MP_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"),
                     ("level", "<i4"), ("track_in_view", "<i4"), ("bad", "<i4")])


class Bounds(C.Structure):
    _fields_ = [("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float)]


def build_oracle(quiet: bool = True) -> None:
    """Compile the restatement (and _ref when /root/reference exists).  Building the checker
    is not using it."""
    subprocess.run(["make", "-C", ORACLE_DIR], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


class OracleExtractor:
    """Handle on either oracle library's extractor API (oracle/orb_oracle.h)."""

    def __init__(self, lib, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7, has_stages=True):
        self.lib, self.nlevels, self.nfeatures, self.has_stages = lib, nlevels, nfeatures, has_stages
        lib.oo_create.restype = C.c_void_p
        lib.oo_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        lib.oo_destroy.argtypes = [C.c_void_p]
        lib.oo_extract.restype = C.c_int
        lib.oo_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_void_p]
        lib.oo_pyramid_level.restype = C.c_int
        lib.oo_pyramid_level.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_int),
                                         C.POINTER(C.c_int), C.POINTER(C.c_size_t)]
        lib.oo_scale_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        self.h = lib.oo_create(nfeatures, scale, nlevels, ini_th, min_th)
        assert self.h

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.oo_destroy(self.h)
            self.h = None

    def extract(self, img: np.ndarray):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        cap = self.nfeatures + 4 * self.nlevels + 16
        kps = np.zeros(cap, dtype=KP_DTYPE)
        desc = np.zeros((cap, 32), dtype=np.uint8)
        counts = np.zeros(self.nlevels, dtype=np.int32)
        n = self.lib.oo_extract(self.h, img.ctypes.data, img.shape[0], img.shape[1], img.strides[0],
                                kps.ctypes.data, desc.ctypes.data, cap, counts.ctypes.data)
        assert n >= 0
        return kps[:n].copy(), desc[:n].copy(), counts

    def pyramid_level(self, level: int, border: int = 19) -> np.ndarray:
        """Bordered level buffer ((h+38) x (w+38)) of the last extract call."""
        data, w, h, step = C.c_void_p(), C.c_int(), C.c_int(), C.c_size_t()
        rc = self.lib.oo_pyramid_level(self.h, level, C.byref(data), C.byref(w), C.byref(h), C.byref(step))
        assert rc == 0
        base = data.value - border * step.value - border
        rows = h.value + 2 * border
        buf = (C.c_uint8 * (rows * step.value)).from_address(base)
        a = np.frombuffer(buf, dtype=np.uint8).reshape(rows, step.value)[:, : w.value + 2 * border]
        return a.copy()

    def scale_tables(self):
        out = [np.zeros(self.nlevels, dtype=np.float32) for _ in range(4)]
        self.lib.oo_scale_tables(self.h, *[o.ctypes.data for o in out])
        return out

    def features_per_level(self):
        out = np.zeros(self.nlevels, dtype=np.int32)
        self.lib.oo_features_per_level.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.oo_features_per_level(self.h, out.ctypes.data)
        return out

    # stage taps (restatement only)
    def candidates(self, level: int):
        cap = 1 << 17
        x, y, s = (np.zeros(cap, dtype=np.int32) for _ in range(3))
        f = self.lib.oo_stage_candidates
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        n = f(self.h, level, x.ctypes.data, y.ctypes.data, s.ctypes.data, cap)
        assert n <= cap
        return x[:n].copy(), y[:n].copy(), s[:n].copy()

    def blurred(self, level: int, w: int, h: int):
        out = np.zeros((h, w), dtype=np.uint8)
        f = self.lib.oo_stage_blurred
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        return out if f(self.h, level, out.ctypes.data) else None

    def level_keypoints(self, level: int):
        cap = self.nfeatures + 64
        x, y = (np.zeros(cap, dtype=np.int32) for _ in range(2))
        r, a = (np.zeros(cap, dtype=np.float32) for _ in range(2))
        f = self.lib.oo_stage_level_keypoints
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int]
        n = f(self.h, level, x.ctypes.data, y.ctypes.data, r.ctypes.data, a.ctypes.data, cap)
        return x[:n].copy(), y[:n].copy(), r[:n].copy(), a[:n].copy()


_libs = {}


def load(kind: str = "port"):
    """kind: 'port' (restatement), 'ref' (verbatim ORBextractor.cc build) or 'mref' (verbatim ORBmatcher.cc build);
    the verbatim builds are None where oracle/_ref was not built."""
    if kind in _libs:
        return _libs[kind]
    path = os.path.join(ORACLE_DIR, {"port": "liborb_oracle.so", "ref": "_ref/liborb_ref.so",
                                     "mref": "_ref/libmatcher_ref.so", "fref": "_ref/libframe_ref.so",
                                     # NOT an oracle: the PRODUCT's drop-in translation units (multi_orb_slam_b200/dropin/)
                                     # behind the same C harness as the verbatim reference builds (tests/native/Makefile)
                                     "dropin": "../tests/native/_build/libmatcher_dropin.so",
                                     "xdropin": "../tests/native/_build/libextractor_dropin.so"}[kind])
    if kind == "port" and not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path) if os.path.exists(path) else None
    _libs[kind] = lib
    return lib


def extractor(kind="port", **kw):
    lib = load(kind)
    if lib is None:
        return None
    return OracleExtractor(lib, has_stages=(kind == "port"), **kw)


# ---- matcher oracle -------------------------------------------------------------------------

def distance(a: np.ndarray, b: np.ndarray) -> int:
    lib = load("port")
    lib.om_distance.restype = C.c_int
    lib.om_distance.argtypes = [C.c_void_p, C.c_void_p]
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    return lib.om_distance(a.ctypes.data, b.ctypes.data)


def bruteforce(q: np.ndarray, t: np.ndarray, ratio: float = 0.9, th: int = 50):
    lib = load("port")
    q = np.ascontiguousarray(q, dtype=np.uint8)
    t = np.ascontiguousarray(t, dtype=np.uint8)
    nq, nt = q.shape[0], t.shape[0]
    idx, d1, d2 = (np.zeros(nq, dtype=np.int32) for _ in range(3))
    lib.om_bruteforce.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p]
    lib.om_bruteforce(q.ctypes.data, nq, t.ctypes.data, nt, ratio, th, idx.ctypes.data, d1.ctypes.data,
                      d2.ctypes.data)
    return idx, d1, d2


def features_in_area(kx, ky, koct, bounds, x, y, r, min_level, max_level):
    lib = load("port")
    kx = np.ascontiguousarray(kx, dtype=np.float32)
    ky = np.ascontiguousarray(ky, dtype=np.float32)
    koct = np.ascontiguousarray(koct, dtype=np.int32)
    out = np.zeros(len(kx) + 1, dtype=np.int32)
    f = lib.om_features_in_area
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, Bounds, C.c_float, C.c_float, C.c_float,
                  C.c_int, C.c_int, C.c_void_p, C.c_int]
    n = f(kx.ctypes.data, ky.ctypes.data, koct.ctypes.data, len(kx), Bounds(*bounds), x, y, r, min_level,
          max_level, out.ctypes.data, len(out))
    return out[:n].copy()


def search_for_initialization(k1, d1, k2, d2, bounds2, prev_xy, window=100, nnratio=0.9, check_ori=True, impl="port"):
    lib = load({"port": "port", "ref": "mref", "mref": "mref", "dropin": "dropin"}[impl])
    k1 = np.ascontiguousarray(k1, dtype=KP_DTYPE)
    k2 = np.ascontiguousarray(k2, dtype=KP_DTYPE)
    d1 = np.ascontiguousarray(d1, dtype=np.uint8)
    d2 = np.ascontiguousarray(d2, dtype=np.uint8)
    prev = np.ascontiguousarray(prev_xy, dtype=np.float32).copy()
    m12 = np.zeros(len(k1), dtype=np.int32)
    f = getattr(lib, ("om_" if impl == "port" else "omr_") + "search_for_initialization")
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, Bounds, C.c_void_p,
                  C.c_int, C.c_float, C.c_int, C.c_void_p]
    n = f(k1.ctypes.data, d1.ctypes.data, len(k1), k2.ctypes.data, d2.ctypes.data, len(k2), Bounds(*bounds2),
          prev.ctypes.data, window, nnratio, int(check_ori), m12.ctypes.data)
    return n, m12, prev


def search_by_projection_points(k, d, u_right, bounds, scale_factors, mp, mp_desc, mp_obs, th, nnratio,
                                frame_mp=None, frame_mp_obs=None, impl="port"):
    lib = load({"port": "port", "ref": "mref", "mref": "mref", "dropin": "dropin"}[impl])
    k = np.ascontiguousarray(k, dtype=KP_DTYPE)
    d = np.ascontiguousarray(d, dtype=np.uint8)
    n = len(k)
    u_right = None if u_right is None else np.ascontiguousarray(u_right, dtype=np.float32)  # NULL = monocular frame
    sf = np.ascontiguousarray(scale_factors, dtype=np.float32)
    mp = np.ascontiguousarray(mp, dtype=MP_DTYPE)
    mp_desc = np.ascontiguousarray(mp_desc, dtype=np.uint8)
    mp_obs = np.ascontiguousarray(mp_obs, dtype=np.int32)
    fmp = np.full(n, -1, dtype=np.int32) if frame_mp is None else np.ascontiguousarray(frame_mp, dtype=np.int32).copy()
    fobs = np.zeros(n, dtype=np.int32) if frame_mp_obs is None else np.ascontiguousarray(frame_mp_obs, dtype=np.int32)
    f = getattr(lib, ("om_" if impl == "port" else "omr_") + "search_by_projection_points")
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, Bounds, C.c_void_p, C.c_int, C.c_void_p,
                  C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    nm = f(k.ctypes.data, d.ctypes.data, None if u_right is None else u_right.ctypes.data, n, Bounds(*bounds), sf.ctypes.data,
           len(sf), mp.ctypes.data, mp_desc.ctypes.data, mp_obs.ctypes.data, len(mp), th, nnratio, fmp.ctypes.data,
           fobs.ctypes.data)
    return nm, fmp


class Camera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("mb", C.c_float),
                ("mbf", C.c_float)]


def search_by_projection_frame(cur_k, cur_desc, cur_uright, cur_cam, bounds, sf, cam, Tcw_cur, Tcw_last, last_k, last_cam,
                               last_valid, last_xyz, last_desc, last_obs, calib, th, mono, check_ori, cur_mp, cur_mp_obs, impl="port"):
    lib = load({"port": "port", "ref": "mref", "mref": "mref", "dropin": "dropin"}[impl])
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    cur_k, last_k = np.ascontiguousarray(cur_k, dtype=KP_DTYPE), np.ascontiguousarray(last_k, dtype=KP_DTYPE)
    cur_desc, last_desc = np.ascontiguousarray(cur_desc, dtype=np.uint8), np.ascontiguousarray(last_desc, dtype=np.uint8)
    ur, sf, tc, tl, xyz, cal = f32(cur_uright), f32(sf), f32(Tcw_cur), f32(Tcw_last), f32(last_xyz), f32(calib)
    ccam, lcam, lval, lobs, fobs = i32(cur_cam), i32(last_cam), i32(last_valid), i32(last_obs), i32(cur_mp_obs)
    fmp = i32(cur_mp).copy()
    f = getattr(lib, ("om_" if impl == "port" else "omr_") + "search_by_projection_frame")
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 4 + [C.c_int, Bounds, C.c_void_p, C.c_int, Camera] + [C.c_void_p] * 8 + \
                 [C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    nm = f(cur_k.ctypes.data, cur_desc.ctypes.data, ur.ctypes.data, ccam.ctypes.data, len(cur_k), Bounds(*bounds),
           sf.ctypes.data, len(sf), Camera(*cam), tc.ctypes.data, tl.ctypes.data, last_k.ctypes.data, lcam.ctypes.data,
           lval.ctypes.data, xyz.ctypes.data, last_desc.ctypes.data, lobs.ctypes.data, len(last_k), cal.ctypes.data, th,
           int(mono), int(check_ori), fmp.ctypes.data, fobs.ctypes.data)
    return nm, fmp


def search_by_projection_keyframe(cur_k, cur_desc, bounds, sf, log_sf, cam, Tcw_cur, kf_valid, kf_xyz, kf_max_dist,
                                  kf_min_dist, kf_max_d, kf_angle, kf_desc, th, orb_dist, check_ori, cur_mp, impl="port"):
    lib = load({"port": "port", "ref": "mref", "mref": "mref", "dropin": "dropin"}[impl])
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    cur_k = np.ascontiguousarray(cur_k, dtype=KP_DTYPE)
    cur_desc, kf_desc = np.ascontiguousarray(cur_desc, dtype=np.uint8), np.ascontiguousarray(kf_desc, dtype=np.uint8)
    sf, tc, xyz, mx, mn, md, ang = f32(sf), f32(Tcw_cur), f32(kf_xyz), f32(kf_max_dist), f32(kf_min_dist), f32(kf_max_d), f32(kf_angle)
    val = np.ascontiguousarray(kf_valid, dtype=np.int32)
    fmp = np.ascontiguousarray(cur_mp, dtype=np.int32).copy()
    f = getattr(lib, ("om_" if impl == "port" else "omr_") + "search_by_projection_keyframe")
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, Bounds, C.c_void_p, C.c_int, C.c_float, Camera] + [C.c_void_p] * 8 + \
                 [C.c_int, C.c_float, C.c_int, C.c_int, C.c_void_p]
    nm = f(cur_k.ctypes.data, cur_desc.ctypes.data, len(cur_k), Bounds(*bounds), sf.ctypes.data, len(sf), log_sf,
           Camera(*cam), tc.ctypes.data, val.ctypes.data, xyz.ctypes.data, mx.ctypes.data, mn.ctypes.data, md.ctypes.data,
           ang.ctypes.data, kf_desc.ctypes.data, len(val), th, orb_dist, int(check_ori), fmp.ctypes.data)
    return nm, fmp


def search_by_projection_sim3(kf_k, kf_desc, kf_cam, bounds, sf, log_sf, cam, Scw, calib, mp_valid, mp_xyz, mp_normal,
                              mp_max_dist, mp_min_dist, mp_max_d, mp_desc, th, matched, impl="port"):
    lib = load({"port": "port", "ref": "mref", "mref": "mref", "dropin": "dropin"}[impl])
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    kf_k = np.ascontiguousarray(kf_k, dtype=KP_DTYPE)
    kf_desc, mp_desc = np.ascontiguousarray(kf_desc, dtype=np.uint8), np.ascontiguousarray(mp_desc, dtype=np.uint8)
    kcam, val = np.ascontiguousarray(kf_cam, dtype=np.int32), np.ascontiguousarray(mp_valid, dtype=np.int32)
    sf, S, cal, xyz, nrm, mx, mn, md = (f32(a) for a in (sf, Scw, calib, mp_xyz, mp_normal, mp_max_dist, mp_min_dist, mp_max_d))
    out = np.ascontiguousarray(matched, dtype=np.int32).copy()
    f = getattr(lib, ("om_" if impl == "port" else "omr_") + "search_by_projection_sim3")
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, Bounds, C.c_void_p, C.c_int, C.c_float, Camera] + \
                 [C.c_void_p] * 9 + [C.c_int, C.c_int, C.c_void_p]
    nm = f(kf_k.ctypes.data, kf_desc.ctypes.data, kcam.ctypes.data, len(kf_k), Bounds(*bounds), sf.ctypes.data, len(sf), log_sf,
           Camera(*cam), S.ctypes.data, cal.ctypes.data, val.ctypes.data, xyz.ctypes.data, nrm.ctypes.data, mx.ctypes.data,
           mn.ctypes.data, md.ctypes.data, mp_desc.ctypes.data, len(val), int(th), out.ctypes.data)
    return nm, out


def fuse(kf_k, kf_desc, kf_uright, kf_cam, bounds, sf, inv_sigma2, log_sf, cam, Tcw, Ow, calib, mp_valid, mp_xyz, mp_normal,
         mp_max_dist, mp_min_dist, mp_max_d, mp_desc, th):
    lib = load("port")
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    kf_k = np.ascontiguousarray(kf_k, dtype=KP_DTYPE)
    kf_desc, mp_desc = np.ascontiguousarray(kf_desc, dtype=np.uint8), np.ascontiguousarray(mp_desc, dtype=np.uint8)
    kcam, val = np.ascontiguousarray(kf_cam, dtype=np.int32), np.ascontiguousarray(mp_valid, dtype=np.int32)
    ur, sf, isg, T, O, cal, xyz, nrm, mx, mn, md = (f32(a) for a in (kf_uright, sf, inv_sigma2, Tcw, Ow, calib, mp_xyz, mp_normal,
                                                                    mp_max_dist, mp_min_dist, mp_max_d))
    best = np.empty((len(val), 2), dtype=np.int32)
    f = lib.om_fuse
    f.restype = C.c_int
    f.argtypes = [C.c_void_p] * 4 + [C.c_int, Bounds, C.c_void_p, C.c_void_p, C.c_int, C.c_float, Camera] + [C.c_void_p] * 10 + \
                 [C.c_int, C.c_float, C.c_void_p]
    n = f(kf_k.ctypes.data, kf_desc.ctypes.data, ur.ctypes.data, kcam.ctypes.data, len(kf_k), Bounds(*bounds), sf.ctypes.data,
          isg.ctypes.data, len(sf), log_sf, Camera(*cam), T.ctypes.data, O.ctypes.data, cal.ctypes.data, val.ctypes.data,
          xyz.ctypes.data, nrm.ctypes.data, mx.ctypes.data, mn.ctypes.data, md.ctypes.data, mp_desc.ctypes.data, len(val),
          float(th), best.ctypes.data)
    return n, best


def fuse_sim3(kf_k, kf_desc, kf_cam, bounds, sf, log_sf, cam, Scw, calib, mp_valid, mp_xyz, mp_normal, mp_max_dist, mp_min_dist,
              mp_max_d, mp_desc, th):
    lib = load("port")
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    kf_k = np.ascontiguousarray(kf_k, dtype=KP_DTYPE)
    kf_desc, mp_desc = np.ascontiguousarray(kf_desc, dtype=np.uint8), np.ascontiguousarray(mp_desc, dtype=np.uint8)
    kcam, val = np.ascontiguousarray(kf_cam, dtype=np.int32), np.ascontiguousarray(mp_valid, dtype=np.int32)
    sf, S, cal, xyz, nrm, mx, mn, md = (f32(a) for a in (sf, Scw, calib, mp_xyz, mp_normal, mp_max_dist, mp_min_dist, mp_max_d))
    best = np.empty((len(val), 2), dtype=np.int32)
    f = lib.om_fuse_sim3
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, Bounds, C.c_void_p, C.c_int, C.c_float, Camera] + \
                 [C.c_void_p] * 9 + [C.c_int, C.c_float, C.c_void_p]
    n = f(kf_k.ctypes.data, kf_desc.ctypes.data, kcam.ctypes.data, len(kf_k), Bounds(*bounds), sf.ctypes.data, len(sf), log_sf,
          Camera(*cam), S.ctypes.data, cal.ctypes.data, val.ctypes.data, xyz.ctypes.data, nrm.ctypes.data, mx.ctypes.data,
          mn.ctypes.data, md.ctypes.data, mp_desc.ctypes.data, len(val), float(th), best.ctypes.data)
    return n, best


def search_by_sim3(k1, d1, cam1, T1w, k2, d2, cam2, T2w, bounds, sf, log_sf, cam, s12, R12, t12, calib, mp1, mp2, th, impl="port"):
    lib = load({"port": "port", "ref": "mref", "mref": "mref", "dropin": "dropin"}[impl])
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    k1, k2 = np.ascontiguousarray(k1, dtype=KP_DTYPE), np.ascontiguousarray(k2, dtype=KP_DTYPE)
    d1, d2 = np.ascontiguousarray(d1, dtype=np.uint8), np.ascontiguousarray(d2, dtype=np.uint8)
    c1, c2, sf = i32(cam1), i32(cam2), f32(sf)
    A = [f32(T1w), f32(T2w), f32(R12), f32(t12), f32(calib)]
    P = []
    for mp in (mp1, mp2):
        P += [i32(mp["valid"]), f32(mp["xyz"]), f32(mp["max_dist"]), f32(mp["min_dist"]), f32(mp["max_d"]),
              np.ascontiguousarray(mp["desc"], dtype=np.uint8)]
    m12 = np.empty(len(k1), dtype=np.int32)
    f = getattr(lib, ("om_" if impl == "port" else "omr_") + "search_by_sim3")
    f.restype = C.c_int
    side = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    f.argtypes = side + side + [Bounds, C.c_void_p, C.c_int, C.c_float, Camera, C.c_float] + [C.c_void_p] * 3 + [C.c_void_p] * 12 + \
        [C.c_float, C.c_void_p]
    p = lambda a: a.ctypes.data
    n = f(p(k1), p(d1), p(c1), len(k1), p(A[0]), p(k2), p(d2), p(c2), len(k2), p(A[1]), Bounds(*bounds), p(sf), len(sf), log_sf,
          Camera(*cam), float(s12), p(A[2]), p(A[3]), p(A[4]), *[p(a) for a in P], float(th), p(m12))
    return n, m12


def bow_transform(voc, desc, levelsup=4):
    """voc: dict(child_start, child_ids, node_desc, word_id, node_weight, L) as synth.random_vocabulary returns."""
    lib = load("port")
    cs, ci = np.ascontiguousarray(voc["child_start"], dtype=np.int32), np.ascontiguousarray(voc["child_ids"], dtype=np.int32)
    nd = np.ascontiguousarray(voc["node_desc"], dtype=np.uint8)
    wi, ww = np.ascontiguousarray(voc["word_id"], dtype=np.int32), np.ascontiguousarray(voc["node_weight"], dtype=np.float64)
    d = np.ascontiguousarray(desc, dtype=np.uint8).reshape(-1, 32)
    n = len(d)
    word, node, weight = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float64)
    bw, bv = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.float64)
    fn, fs, fi = np.empty(max(n, 1), np.int32), np.empty(n + 1, np.int32), np.empty(max(n, 1), np.int32)
    nb, nf = C.c_int(0), C.c_int(0)
    f = lib.om_bow_transform
    f.restype = None
    f.argtypes = [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.POINTER(C.c_int)] + \
        [C.c_void_p] * 3 + [C.POINTER(C.c_int)]
    f(cs.ctypes.data, ci.ctypes.data, nd.ctypes.data, wi.ctypes.data, ww.ctypes.data, len(nd), int(voc["L"]), d.ctypes.data, n,
      int(levelsup), word.ctypes.data, node.ctypes.data, weight.ctypes.data, bw.ctypes.data, bv.ctypes.data, C.byref(nb),
      fn.ctypes.data, fs.ctypes.data, fi.ctypes.data, C.byref(nf))
    return dict(word=word, node=node, weight=weight, bow=(bw[: nb.value].copy(), bv[: nb.value].copy()),
                featvec=(fn[: nf.value].copy(), fs[: nf.value + 1].copy(), fi[: fs[nf.value]].copy()))


def compute_distinctive_descriptors(desc, offsets):
    lib = load("port")
    d = np.ascontiguousarray(desc, dtype=np.uint8).reshape(-1, 32)
    off = np.ascontiguousarray(offsets, dtype=np.int32)
    best = np.empty(len(off) - 1, dtype=np.int32)
    lib.om_compute_distinctive_descriptors.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.om_compute_distinctive_descriptors(d.ctypes.data, off.ctypes.data, len(off) - 1, best.ctypes.data)
    return best


def undistort_points(pts, fx, fy, cx, cy, dist5):
    lib = load("port")
    p = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 2)
    d = np.ascontiguousarray(dist5, dtype=np.float32)
    out = np.empty_like(p)
    lib.cvp_undistort_points.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    lib.cvp_undistort_points(p.ctypes.data, len(p), fx, fy, cx, cy, d.ctypes.data, out.ctypes.data)
    return out


def undistort_keypoints(k, fx, fy, cx, cy, dist5):
    lib = load("port")
    k = np.ascontiguousarray(k, dtype=KP_DTYPE)
    d = np.ascontiguousarray(dist5, dtype=np.float32)
    out = np.empty_like(k)
    lib.om_undistort_keypoints.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    lib.om_undistort_keypoints(k.ctypes.data, len(k), fx, fy, cx, cy, d.ctypes.data, out.ctypes.data)
    return out


def compute_image_bounds(cols, rows, fx, fy, cx, cy, dist5):
    lib = load("port")
    d = np.ascontiguousarray(dist5, dtype=np.float32)
    b = Bounds()
    lib.om_compute_image_bounds.argtypes = [C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p,
                                            C.POINTER(Bounds)]
    lib.om_compute_image_bounds(cols, rows, fx, fy, cx, cy, d.ctypes.data, C.byref(b))
    return (b.min_x, b.max_x, b.min_y, b.max_y)


def compute_stereo_from_rgbd(k, k_un, depth, mbf):
    lib = load("port")
    k, k_un = np.ascontiguousarray(k, dtype=KP_DTYPE), np.ascontiguousarray(k_un, dtype=KP_DTYPE)
    depth = np.ascontiguousarray(depth, dtype=np.float32)
    ur, dz = np.empty(len(k), np.float32), np.empty(len(k), np.float32)
    lib.om_compute_stereo_from_rgbd.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_size_t,
                                                C.c_float, C.c_void_p, C.c_void_p]
    lib.om_compute_stereo_from_rgbd(k.ctypes.data, k_un.ctypes.data, len(k), depth.ctypes.data, depth.shape[1], depth.shape[0],
                                    depth.shape[1], mbf, ur.ctypes.data, dz.ctypes.data)
    return ur, dz


class OmImage(C.Structure):
    _fields_ = [("data", C.c_void_p), ("w", C.c_int32), ("h", C.c_int32), ("step", C.c_size_t)]


def compute_stereo_matches(kl, dl, kr, dr, pyr_l, pyr_r, scale, inv_scale, mbf, mb, border=19, impl="port"):
    """pyr_l / pyr_r: per level the BORDERED level image ((h+2*border) x (w+2*border) u8) of each extractor.
    impl "ref" runs the reference's own (un-commented) Frame::ComputeStereoMatches of oracle/_ref/libframe_ref.so."""
    lib = load("port" if impl == "port" else "fref")
    kl, kr = np.ascontiguousarray(kl, dtype=KP_DTYPE), np.ascontiguousarray(kr, dtype=KP_DTYPE)
    dl, dr = np.ascontiguousarray(dl, dtype=np.uint8), np.ascontiguousarray(dr, dtype=np.uint8)
    scale, inv_scale = np.ascontiguousarray(scale, dtype=np.float32), np.ascontiguousarray(inv_scale, dtype=np.float32)
    keep = [[np.ascontiguousarray(a) for a in p] for p in (pyr_l, pyr_r)]
    views = []
    for p in keep:
        arr = (OmImage * len(p))()
        for i, a in enumerate(p):
            arr[i] = OmImage(a.ctypes.data + border * a.strides[0] + border, a.shape[1] - 2 * border, a.shape[0] - 2 * border,
                             a.strides[0])
        views.append(arr)
    uright, depth = np.empty(len(kl), np.float32), np.empty(len(kl), np.float32)
    f = lib.om_compute_stereo_matches if impl == "port" else lib.ofr_compute_stereo_matches
    f.restype = None
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                  C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
    f(kl.ctypes.data, dl.ctypes.data, len(kl), kr.ctypes.data, dr.ctypes.data, len(kr), views[0], views[1], len(pyr_l),
      scale.ctypes.data, inv_scale.ctypes.data, mbf, mb, uright.ctypes.data, depth.ctypes.data)
    return uright, depth


def assign_features_to_grid(k_un, bounds):
    lib = load("port")
    k_un = np.ascontiguousarray(k_un, dtype=KP_DTYPE)
    start, items = np.zeros(64 * 48 + 1, np.int32), np.zeros(max(len(k_un), 1), np.int32)
    lib.om_assign_features_to_grid.argtypes = [C.c_void_p, C.c_int, Bounds, C.c_void_p, C.c_void_p]
    lib.om_assign_features_to_grid(k_un.ctypes.data, len(k_un), Bounds(*bounds), start.ctypes.data, items.ctypes.data)
    return start, items[: start[-1]]


def search_by_bow(d1, angle1, valid1, fv1, d2, angle2, valid2, fv2, nnratio=0.7, check_ori=True, max_dist=50):
    """fv = (node_ids, start, items) CSR feature vector.  Returns (nmatches, matches12, matches21)."""
    lib = load("port")
    d1 = np.ascontiguousarray(d1, dtype=np.uint8)
    d2 = np.ascontiguousarray(d2, dtype=np.uint8)
    a1 = np.ascontiguousarray(angle1, dtype=np.float32)
    a2 = np.ascontiguousarray(angle2, dtype=np.float32)
    v1 = None if valid1 is None else np.ascontiguousarray(valid1, dtype=np.int32)
    v2 = None if valid2 is None else np.ascontiguousarray(valid2, dtype=np.int32)
    f1 = [np.ascontiguousarray(a, dtype=np.int32) for a in fv1]
    f2 = [np.ascontiguousarray(a, dtype=np.int32) for a in fv2]
    n1, n2 = len(d1), len(d2)
    m12, m21 = np.zeros(n1, dtype=np.int32), np.zeros(n2, dtype=np.int32)
    f = lib.om_search_by_bow
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] * 2 + \
        [C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    n = f(d1.ctypes.data, a1.ctypes.data, None if v1 is None else v1.ctypes.data, n1, f1[0].ctypes.data, f1[1].ctypes.data,
          f1[2].ctypes.data, len(f1[0]),
          d2.ctypes.data, a2.ctypes.data, None if v2 is None else v2.ctypes.data, n2, f2[0].ctypes.data, f2[1].ctypes.data,
          f2[2].ctypes.data, len(f2[0]), nnratio, int(check_ori), int(max_dist), m12.ctypes.data, m21.ctypes.data)
    return n, m12, m21


def search_for_triangulation(sc, fv1, fv2, only_stereo=False, cam_enabled=(1, 1), check_ori=True):
    """sc: dict of synth.triangulation_scene arrays.  Returns (nmatches, matches12)."""
    lib = load("port")
    c = lambda a, t: np.ascontiguousarray(a, dtype=t)
    k1, k2 = c(sc["k1"], KP_DTYPE), c(sc["k2"], KP_DTYPE)
    d1, d2 = c(sc["d1"], np.uint8), c(sc["d2"], np.uint8)
    arrs = [k1, d1, c(sc["has_mp1"], np.int32), c(sc["cam1"], np.int32), c(sc["uright1"], np.float32)]
    f1 = [c(a, np.int32) for a in fv1]
    arrs2 = [k2, d2, c(sc["has_mp2"], np.int32), c(sc["cam2"], np.int32), c(sc["uright2"], np.float32)]
    f2 = [c(a, np.int32) for a in fv2]
    tail = [c(sc["F12s"], np.float32), c(sc["epipoles"], np.float32), c(sc["scale_factors"], np.float32),
            c(sc["level_sigma2"], np.float32)]
    en = c(cam_enabled, np.int32)
    m12 = np.zeros(len(k1), dtype=np.int32)
    f = lib.om_search_for_triangulation
    f.restype = C.c_int
    side = [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int]
    f.argtypes = side + side + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
    p = lambda a: a.ctypes.data
    n = f(*[p(a) for a in arrs], len(k1), p(f1[0]), p(f1[1]), p(f1[2]), len(f1[0]),
          *[p(a) for a in arrs2], len(k2), p(f2[0]), p(f2[1]), p(f2[2]), len(f2[0]),
          *[p(a) for a in tail], int(only_stereo), p(en), int(check_ori), p(m12))
    return n, m12


def search_by_bow_ref(variant, d1, angle1, valid1, fv1, d2, angle2, valid2, fv2, nnratio=0.7, check_ori=True):
    """The reference's own SearchByBoW (verbatim build): variant 0 = (KeyFrame*, Frame&), 1 = (KeyFrame*, KeyFrame*)."""
    lib = load("mref")
    c = lambda a, t: np.ascontiguousarray(a, dtype=t)
    d1, d2, a1, a2 = c(d1, np.uint8), c(d2, np.uint8), c(angle1, np.float32), c(angle2, np.float32)
    v1 = None if valid1 is None else c(valid1, np.int32)
    v2 = None if valid2 is None else c(valid2, np.int32)
    f1, f2 = [c(a, np.int32) for a in fv1], [c(a, np.int32) for a in fv2]
    m12, m21 = np.zeros(len(d1), np.int32), np.zeros(len(d2), np.int32)
    f = lib.omr_search_by_bow
    f.restype = C.c_int
    side = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    f.argtypes = [C.c_int] + side + side + [C.c_float, C.c_int, C.c_void_p, C.c_void_p]
    p = lambda a: None if a is None else a.ctypes.data
    n = f(variant, p(d1), p(a1), p(v1), len(d1), p(f1[0]), p(f1[1]), p(f1[2]), len(f1[0]),
          p(d2), p(a2), p(v2), len(d2), p(f2[0]), p(f2[1]), p(f2[2]), len(f2[0]), nnratio, int(check_ori), p(m12), p(m21))
    return n, m12, m21


def search_for_triangulation_ref(sc, fv1, fv2, T1w, T2w, Cw1, cam, only_stereo=False, cam_enabled=(1, 1), check_ori=True):
    """The reference's SearchForTriangulation on poses.  Returns (nmatches, matches12, F12s [2,3,3], epipoles [4])."""
    lib = load("mref")
    c = lambda a, t: np.ascontiguousarray(a, dtype=t)
    k1, k2 = c(sc["k1"], KP_DTYPE), c(sc["k2"], KP_DTYPE)
    a1 = [k1, c(sc["d1"], np.uint8), c(sc["has_mp1"], np.int32), c(sc["cam1"], np.int32), c(sc["uright1"], np.float32)]
    a2 = [k2, c(sc["d2"], np.uint8), c(sc["has_mp2"], np.int32), c(sc["cam2"], np.int32), c(sc["uright2"], np.float32)]
    f1, f2 = [c(a, np.int32) for a in fv1], [c(a, np.int32) for a in fv2]
    T1, T2, Cw = c(T1w, np.float32), c(T2w, np.float32), c(Cw1, np.float32)
    sf, ls, en = c(sc["scale_factors"], np.float32), c(sc["level_sigma2"], np.float32), c(cam_enabled, np.int32)
    m12, F, epi = np.zeros(len(k1), np.int32), np.zeros((2, 3, 3), np.float32), np.zeros(4, np.float32)
    f = lib.omr_search_for_triangulation
    f.restype = C.c_int
    side = [C.c_void_p] * 5 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int]
    f.argtypes = side + side + [C.c_void_p] * 3 + [Camera, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 3
    p = lambda a: a.ctypes.data
    n = f(*[p(a) for a in a1], len(k1), p(f1[0]), p(f1[1]), p(f1[2]), len(f1[0]), *[p(a) for a in a2], len(k2), p(f2[0]), p(f2[1]),
          p(f2[2]), len(f2[0]), p(T1), p(T2), p(Cw), Camera(*cam), p(sf), p(ls), len(sf), int(only_stereo), p(en), int(check_ori),
          p(m12), p(F), p(epi))
    return n, m12, F, epi


def fuse_ref(sim3, kf_k, kf_desc, kf_uright, kf_cam, kf_held, bounds, sf, inv_sigma2, log_sf, cam, pose, Ow, calib, mp_valid, mp_xyz,
             mp_normal, mp_max_dist, mp_min_dist, mp_max_d, mp_desc, th):
    lib = load("mref")
    f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
    i32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)
    kf_k = np.ascontiguousarray(kf_k, dtype=KP_DTYPE)
    kf_desc, mp_desc = np.ascontiguousarray(kf_desc, dtype=np.uint8), np.ascontiguousarray(mp_desc, dtype=np.uint8)
    A = [f32(kf_uright), i32(kf_cam), i32(kf_held)]
    B = [f32(sf), f32(inv_sigma2)]
    Cc = [f32(pose), f32(Ow), f32(calib), i32(mp_valid), f32(mp_xyz), f32(mp_normal), f32(mp_max_dist), f32(mp_min_dist), f32(mp_max_d)]
    best = np.empty((len(Cc[3]), 2), dtype=np.int32)
    f = lib.omr_fuse
    f.restype = C.c_int
    f.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, Bounds, C.c_void_p, C.c_void_p, C.c_int, C.c_float, Camera] + \
        [C.c_void_p] * 10 + [C.c_int, C.c_float, C.c_void_p]
    p = lambda a: None if a is None else a.ctypes.data
    n = f(int(sim3), p(kf_k), p(kf_desc), p(A[0]), p(A[1]), p(A[2]), len(kf_k), Bounds(*bounds), p(B[0]), p(B[1]), len(B[0]), log_sf,
          Camera(*cam), *[p(a) for a in Cc], p(mp_desc), len(Cc[3]), float(th), p(best))
    return n, best


def write_vocabulary_text(voc, path, k=None):
    """The ORBvoc.txt format the reference's loadFromTextFile reads (TemplatedVocabulary.h:1339-1424): header
    `k L scoring weighting` (L1_NORM = 0, TF_IDF = 0), then one line per node in id order:
    `parent isLeaf d0 .. d31 weight`.  Weights with 17 significant digits so that they round-trip exactly."""
    cs, ci = voc["child_start"], voc["child_ids"]
    n = len(voc["node_desc"])
    parent = np.zeros(n, dtype=np.int64)
    for i in range(n):
        parent[ci[cs[i]:cs[i + 1]]] = i
    kk = int(max(np.diff(cs).max(), 2)) if k is None else k
    lines = [f"{kk} {voc['L']} 0 0"]
    for i in range(1, n):
        leaf = int(cs[i] == cs[i + 1])
        lines.append(f"{parent[i]} {leaf} " + " ".join(str(int(b)) for b in voc["node_desc"][i]) + f" {voc['node_weight'][i]:.17g}")
    with open(path, "w") as f:
        f.write("\n".join(lines))  # no trailing newline: the loader's `while(!f.eof())` would read it as one more node


def bow_transform_ref(voc_path, desc, levelsup=4):
    """The reference's own DBoW2 transform (verbatim build).  Returns dict(bow=(words, values), featvec=(nodes, start, items))."""
    lib = load("mref")
    d = np.ascontiguousarray(desc, dtype=np.uint8).reshape(-1, 32)
    n = len(d)
    bw, bv = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.float64)
    fn, fs, fi = np.empty(max(n, 1), np.int32), np.empty(n + 1, np.int32), np.empty(max(n, 1), np.int32)
    nb, nf = C.c_int(0), C.c_int(0)
    f = lib.omr_bow_transform
    f.restype = C.c_int
    f.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p,
                  C.POINTER(C.c_int)]
    rc = f(str(voc_path).encode(), d.ctypes.data, n, int(levelsup), bw.ctypes.data, bv.ctypes.data, C.byref(nb), fn.ctypes.data,
           fs.ctypes.data, fi.ctypes.data, C.byref(nf))
    assert rc > 0, "the reference could not load the vocabulary file"
    return dict(n_words=rc, bow=(bw[: nb.value].copy(), bv[: nb.value].copy()),
                featvec=(fn[: nf.value].copy(), fs[: nf.value + 1].copy(), fi[: fs[nf.value]].copy()))


def frame_glue_ref(k0, d0, k1, d1, depth0, depth1, cols, rows, fx, fy, cx, cy, dist, bf, nlevels=8, scale_factor=1.2):
    """The reference's two-camera RGB-D Frame constructor (verbatim Frame.cc) on flat inputs.  Returns a dict."""
    lib = load("fref")
    c = lambda a, t: np.ascontiguousarray(a, dtype=t)
    k0, k1, d0, d1 = c(k0, KP_DTYPE), c(k1, KP_DTYPE), c(d0, np.uint8), c(d1, np.uint8)
    z0, z1 = c(depth0, np.float32), c(depth1, np.float32)
    dist = c(dist, np.float32)
    n = len(k0) + len(k1)
    k_un, ur, dz = np.empty(n, KP_DTYPE), np.empty(n, np.float32), np.empty(n, np.float32)
    cam, loc = np.empty(n, np.int32), np.empty(n, np.int32)
    b = Bounds()
    gs, gi = np.zeros((2, 64 * 48 + 1), np.int32), np.zeros((2, max(n, 1)), np.int32)
    f = lib.ofr_frame_glue
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                  C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_float, C.c_void_p,
                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(Bounds), C.c_void_p, C.c_void_p]
    p = lambda a: a.ctypes.data
    rc = f(p(k0), p(d0), len(k0), p(k1), p(d1), len(k1), p(z0), p(z1), cols, rows, fx, fy, cx, cy, p(dist), len(dist), bf, nlevels,
           scale_factor, p(k_un), p(ur), p(dz), p(cam), p(loc), C.byref(b), p(gs), p(gi))
    assert rc == n, rc
    return dict(k_un=k_un, uright=ur, depth=dz, cam=cam, local=loc, bounds=(b.min_x, b.max_x, b.min_y, b.max_y), grid_start=gs,
                grid_items=gi)


def features_in_area_ref(k0, k1, cols, rows, fx, fy, cx, cy, dist, cam, x, y, r, min_level, max_level):
    lib = load("fref")
    k0, k1 = np.ascontiguousarray(k0, dtype=KP_DTYPE), np.ascontiguousarray(k1, dtype=KP_DTYPE)
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    out = np.empty(len(k0) + len(k1) + 1, np.int32)
    f = lib.ofr_features_in_area
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p,
                  C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_int, C.c_void_p, C.c_int]
    n = f(k0.ctypes.data, len(k0), k1.ctypes.data, len(k1), cols, rows, fx, fy, cx, cy, dist.ctypes.data, len(dist), cam, x, y, r,
          min_level, max_level, out.ctypes.data, len(out))
    assert n >= 0
    return out[:n].copy()


def search_for_initialization_frame_ref(k1, d1, k2, d2, cols, rows, fx, fy, cx, cy, dist, prev_xy, window=100, nnratio=0.9,
                                        check_ori=True):
    """SearchForInitialization of the verbatim ORBmatcher.cc over frames built by the verbatim Frame.cc (distorted camera)."""
    lib = load("fref")
    k1, k2 = np.ascontiguousarray(k1, dtype=KP_DTYPE), np.ascontiguousarray(k2, dtype=KP_DTYPE)
    d1, d2 = np.ascontiguousarray(d1, dtype=np.uint8), np.ascontiguousarray(d2, dtype=np.uint8)
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    prev = np.ascontiguousarray(prev_xy, dtype=np.float32).copy()
    m12 = np.zeros(len(k1), dtype=np.int32)
    f = lib.ofr_search_for_initialization
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float,
                  C.c_float, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_void_p]
    n = f(k1.ctypes.data, d1.ctypes.data, len(k1), k2.ctypes.data, d2.ctypes.data, len(k2), cols, rows, fx, fy, cx, cy,
          dist.ctypes.data, len(dist), prev.ctypes.data, window, nnratio, int(check_ori), m12.ctypes.data)
    assert n >= 0
    return n, m12, prev


def distance_ref(a, b, impl="mref"):
    lib = load(impl)
    a, b = np.ascontiguousarray(a, dtype=np.uint8), np.ascontiguousarray(b, dtype=np.uint8)
    lib.omr_distance.argtypes = [C.c_void_p, C.c_void_p]
    return lib.omr_distance(a.ctypes.data, b.ctypes.data)


def three_maxima(counts):
    lib = load("port")
    c = np.ascontiguousarray(counts, dtype=np.int32)
    i1, i2, i3 = C.c_int(), C.c_int(), C.c_int()
    lib.om_three_maxima.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.om_three_maxima(c.ctypes.data, len(c), C.byref(i1), C.byref(i2), C.byref(i3))
    return i1.value, i2.value, i3.value
