"""Host-side mirror of the hot members of the reference's `ORB_SLAM2::ORBmatcher`
(include/ORBmatcher.h:37-137, src/ORBmatcher.cc) over the C-ABI CUDA library, plus the flat
`Frame` / `MapPoints` containers that stand in for the reference's object graphs
(include/Frame.h, include/MapPoint.h): only the fields the matcher reads."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, MP_DTYPE, Bounds, Camera, check_m, lib


@dataclass
class Frame:
    """Fields of ORB_SLAM2::Frame read by the matcher: mvKeysUn, mDescriptors, mvuRight,
    mvScaleFactors, mnMinX/MaxX/MinY/MaxY (image bounds, src/Frame.cc:262-278), mvpMapPoints
    (here: index into the MapPoints passed to the search, -1 = none)."""
    mvKeysUn: np.ndarray
    mDescriptors: np.ndarray
    width: float
    height: float
    mvScaleFactors: Optional[np.ndarray] = None
    mvuRight: Optional[np.ndarray] = None
    mvpMapPoints: Optional[np.ndarray] = None
    mvpMapPointsObserved: Optional[np.ndarray] = None  # Observations()>0 of the held points

    def __post_init__(self):
        self.mvKeysUn = np.ascontiguousarray(self.mvKeysUn, dtype=KP_DTYPE)
        self.mDescriptors = np.ascontiguousarray(self.mDescriptors, dtype=np.uint8).reshape(-1, 32)
        if self.mvpMapPoints is None:
            self.mvpMapPoints = np.full(len(self.mvKeysUn), -1, dtype=np.int32)

    @property
    def N(self) -> int:
        return len(self.mvKeysUn)

    @property
    def bounds(self) -> Bounds:
        return Bounds(0.0, float(self.width), 0.0, float(self.height))


@dataclass
class MapPoints:
    """Struct-of-arrays view of vector<MapPoint*>: the tracking fields written by
    Frame::isInFrustum (src/Frame.cc:443-499) and the representative descriptors."""
    fields: np.ndarray                 # MP_DTYPE
    descriptors: np.ndarray            # [n, 32] u8   (MapPoint::GetDescriptor)
    observed: Optional[np.ndarray] = None  # Observations()>0, default all true

    def __post_init__(self):
        self.fields = np.ascontiguousarray(self.fields, dtype=MP_DTYPE)
        self.descriptors = np.ascontiguousarray(self.descriptors, dtype=np.uint8).reshape(-1, 32)


class ORBmatcher:
    TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30  # src/ORBmatcher.cc:37-39

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True, *, device: int = -1):
        self.mfNNratio, self.mbCheckOrientation = float(nnratio), bool(checkOri)
        h = C.c_void_p()
        rc = lib.orbm_create(device, C.byref(h))
        if rc != _lib.OK:
            raise _lib.OrbError(rc, (lib.orbm_last_error(None) or b"").decode())
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            lib.orbm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return lib.orbm_launch_count(self._h)

    def sync(self):
        check_m(self._h, lib.orbm_sync(self._h))

    def set_stream(self, cuda_stream: int) -> None:
        check_m(self._h, lib.orbm_set_stream(self._h, cuda_stream or None))

    # -- DescriptorDistance (src/ORBmatcher.cc:3994-4010) ---------------------------------------
    def DescriptorDistance(self, a: np.ndarray, b: np.ndarray) -> int:
        return int(self.distance_pairs(np.asarray(a).reshape(1, 32), np.asarray(b).reshape(1, 32))[0])

    def distance_pairs(self, a: np.ndarray, b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(a, dtype=np.uint8).reshape(-1, 32)
        b = np.ascontiguousarray(b, dtype=np.uint8).reshape(-1, 32)
        assert a.shape == b.shape
        out = np.zeros(len(a), dtype=np.int32)
        check_m(self._h, lib.orbm_distance_pairs_host(self._h, a.ctypes.data, b.ctypes.data, len(a), out.ctypes.data))
        return out

    # -- brute force best/second-best with ratio test ---------------------------------------------
    def bruteforce(self, q: np.ndarray, t: np.ndarray, th_dist: int = TH_LOW, ratio: Optional[float] = None):
        q = np.ascontiguousarray(q, dtype=np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(t, dtype=np.uint8).reshape(-1, 32)
        idx, d1, d2 = (np.zeros(len(q), dtype=np.int32) for _ in range(3))
        check_m(self._h, lib.orbm_bruteforce_host(self._h, q.ctypes.data, len(q), t.ctypes.data, len(t),
                                                  self.mfNNratio if ratio is None else ratio, th_dist, idx.ctypes.data,
                                                  d1.ctypes.data, d2.ctypes.data))
        return idx, d1, d2

    def bruteforce_device(self, q, t, idx, d1, d2, th_dist: int = TH_LOW, ratio: Optional[float] = None):
        """torch CUDA tensors: q [nq,32] u8, t [nt,32] u8, idx/d1/d2 [nq] i32.  Asynchronous."""
        check_m(self._h, lib.orbm_bruteforce_device(self._h, q.data_ptr(), q.shape[0], t.data_ptr(), t.shape[0],
                                                    self.mfNNratio if ratio is None else ratio, th_dist, idx.data_ptr(),
                                                    d1.data_ptr(), d2.data_ptr()))

    def bruteforce_batch_device(self, q, nq, t, nt, idx, d1, d2, th_dist: int = TH_LOW, ratio: Optional[float] = None):
        """Many small pairs at once (torch CUDA tensors): q/t [P, cap, 32] u8 (may be strided views along
        dim 0), nq/nt [P] i32, idx/d1/d2 [P, cap] i32.  Asynchronous."""
        P, cap = q.shape[0], q.shape[1]
        assert t.shape[1] == cap and q.stride(1) == 32 and t.stride(1) == 32
        check_m(self._h, lib.orbm_bruteforce_batch_device(
            self._h, P, cap, q.data_ptr(), nq.data_ptr(), q.stride(0), t.data_ptr(), nt.data_ptr(), t.stride(0),
            self.mfNNratio if ratio is None else ratio, th_dist, idx.data_ptr(), d1.data_ptr(), d2.data_ptr()))

    def bruteforce_indexed_device(self, base, tables, idx, d1, d2, cap: int, th_dist: int = TH_LOW,
                                  ratio: Optional[float] = None):
        """Pairs addressed by byte offsets inside one device buffer (dist.RigLayout.match_tables): base uint8 CUDA
        tensor, tables int64 CUDA tensor [4, P] (query rows, target rows, query count, target count), idx/d1/d2
        [P, cap] i32.  One launch for all pairs.  Asynchronous."""
        P = tables.shape[1]
        assert tables.shape[0] == 4 and tables.is_contiguous() and idx.shape[0] >= P and idx.shape[1] == cap
        check_m(self._h, lib.orbm_bruteforce_indexed_device(
            self._h, P, cap, base.data_ptr(), tables[0].data_ptr(), tables[1].data_ptr(), tables[2].data_ptr(),
            tables[3].data_ptr(), self.mfNNratio if ratio is None else ratio, th_dist, idx.data_ptr(), d1.data_ptr(),
            d2.data_ptr()))

    def debug_set_row_budget(self, rows_per_query: int) -> None:
        """Test tap (orbm_debug_set_row_budget): first guess of the candidate rows per query of the single-call
        projection searches; 1 forces the overflow-and-repeat path."""
        check_m(self._h, lib.orbm_debug_set_row_budget(self._h, int(rows_per_query)))

    # -- SearchForInitialization (src/ORBmatcher.cc:868-983) --------------------------------------
    def SearchForInitialization(self, F1: Frame, F2: Frame, vbPrevMatched: np.ndarray, windowSize: int = 10):
        """Returns (nmatches, vnMatches12); vbPrevMatched ([N1,2] f32) is updated in place."""
        n, m12, prev = self.search_for_initialization_batch(
            F1.mvKeysUn[None], F1.mDescriptors[None], np.array([F1.N], dtype=np.int32),
            F2.mvKeysUn[None], F2.mDescriptors[None], np.array([F2.N], dtype=np.int32), F2.bounds,
            np.asarray(vbPrevMatched, dtype=np.float32).reshape(1, -1, 2), windowSize)
        vbPrevMatched[...] = prev[0, : len(vbPrevMatched)]
        return int(n[0]), m12[0, : F1.N].copy()

    def search_for_initialization_batch(self, k1, d1, n1, k2, d2, n2, bounds2: Bounds, prev_xy, windowSize: int = 10):
        """Batch of independent frame pairs.  k1/k2 [P, cap] KP_DTYPE, d1/d2 [P, cap, 32], n1/n2 [P],
        prev_xy [P, cap, 2].  Arrays narrower than the common cap are padded."""
        P = len(n1)
        cap = max(k1.shape[1], k2.shape[1], 1)

        def pad(a, shape, dtype):
            out = np.zeros(shape, dtype=dtype)
            out[:, : a.shape[1]] = a
            return out
        k1p, k2p = pad(k1, (P, cap), KP_DTYPE), pad(k2, (P, cap), KP_DTYPE)
        d1p, d2p = pad(d1, (P, cap, 32), np.uint8), pad(d2, (P, cap, 32), np.uint8)
        prev = pad(np.asarray(prev_xy, dtype=np.float32), (P, cap, 2), np.float32)
        n1 = np.ascontiguousarray(n1, dtype=np.int32)
        n2 = np.ascontiguousarray(n2, dtype=np.int32)
        m12 = np.full((P, cap), -1, dtype=np.int32)
        nm = np.zeros(P, dtype=np.int32)
        check_m(self._h, lib.orbm_search_for_initialization_host(
            self._h, P, cap, k1p.ctypes.data, d1p.ctypes.data, n1.ctypes.data, k2p.ctypes.data, d2p.ctypes.data,
            n2.ctypes.data, bounds2, prev.ctypes.data, int(windowSize), self.mfNNratio, int(self.mbCheckOrientation),
            m12.ctypes.data, nm.ctypes.data))
        return nm, m12, prev

    def search_for_initialization_device(self, P, cap, k1, d1, n1, k2, d2, n2, bounds2: Bounds, prev_xy, windowSize,
                                         matches12, nmatches):
        """Device-resident batch (torch CUDA tensors laid out as the extractor's batch outputs)."""
        check_m(self._h, lib.orbm_search_for_initialization_device(
            self._h, P, cap, k1.data_ptr(), d1.data_ptr(), n1.data_ptr(), k2.data_ptr(), d2.data_ptr(), n2.data_ptr(),
            bounds2, None if prev_xy is None else prev_xy.data_ptr(), int(windowSize), self.mfNNratio, int(self.mbCheckOrientation),
            matches12.data_ptr(), nmatches.data_ptr()))

    # -- SearchByProjection(Frame&, vector<MapPoint*>&, th) (src/ORBmatcher.cc:62-157) ------------
    def SearchByProjection(self, F: Frame, vpMapPoints: MapPoints, th: float = 3.0) -> int:
        """Assigns map-point indices into F.mvpMapPoints; returns nmatches."""
        n = F.N
        sf = np.ascontiguousarray(F.mvScaleFactors, dtype=np.float32)
        ur = None if F.mvuRight is None else np.ascontiguousarray(F.mvuRight, dtype=np.float32)
        obs = None if vpMapPoints.observed is None else np.ascontiguousarray(vpMapPoints.observed, dtype=np.int32)
        fobs = None if F.mvpMapPointsObserved is None else np.ascontiguousarray(F.mvpMapPointsObserved, dtype=np.int32)
        fmp = np.ascontiguousarray(F.mvpMapPoints, dtype=np.int32)
        nm = C.c_int(0)
        check_m(self._h, lib.orbm_search_by_projection_points_host(
            self._h, F.mvKeysUn.ctypes.data, F.mDescriptors.ctypes.data, None if ur is None else ur.ctypes.data, n,
            F.bounds, sf.ctypes.data, len(sf), vpMapPoints.fields.ctypes.data, vpMapPoints.descriptors.ctypes.data,
            None if obs is None else obs.ctypes.data, len(vpMapPoints.fields), float(th), self.mfNNratio,
            fmp.ctypes.data, None if fobs is None else fobs.ctypes.data, C.byref(nm)))
        F.mvpMapPoints = fmp
        return nm.value

    # -- SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, th, bMono, CalibMatrix) ---
    def SearchByProjectionFrame(self, cur: Frame, cur_cam, camera: Camera, Tcw_cur, Tcw_last, last_keys, last_cam,
                                last_valid, last_xyz, last_desc, last_obs, calib, th: float, bMono: bool = False) -> int:
        """src/ORBmatcher.cc:3448-3641 over flat arrays.  `cur` holds the concatenated keypoints of
        all cameras (mvKeysUn_total / descriptors per global index / mvuRight_total); cur.mvpMapPoints
        is updated in place with indices into the last-frame arrays.  Returns nmatches."""
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        i32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)
        n, nl = cur.N, len(last_keys)
        sf = f32(cur.mvScaleFactors)
        ur = None if cur.mvuRight is None else f32(cur.mvuRight)
        ccam, lcam, lval, lobs = i32(cur_cam), i32(last_cam), i32(last_valid), i32(last_obs)
        fobs = i32(cur.mvpMapPointsObserved)
        fmp = np.ascontiguousarray(cur.mvpMapPoints, dtype=np.int32)
        lk = np.ascontiguousarray(last_keys, dtype=KP_DTYPE)
        lxyz, ldesc = f32(last_xyz), np.ascontiguousarray(last_desc, dtype=np.uint8)
        tc, tl, cal = f32(Tcw_cur), f32(Tcw_last), f32(calib)
        p = lambda a: None if a is None else a.ctypes.data
        nm = C.c_int(0)
        check_m(self._h, lib.orbm_search_by_projection_frame_host(
            self._h, cur.mvKeysUn.ctypes.data, cur.mDescriptors.ctypes.data, p(ur), p(ccam), n, cur.bounds, sf.ctypes.data,
            len(sf), camera, tc.ctypes.data, tl.ctypes.data, lk.ctypes.data, p(lcam), lval.ctypes.data, lxyz.ctypes.data,
            ldesc.ctypes.data, p(lobs), nl, cal.ctypes.data, float(th), int(bMono), int(self.mbCheckOrientation),
            fmp.ctypes.data, p(fobs), C.byref(nm)))
        cur.mvpMapPoints = fmp
        return nm.value

    # -- SearchByProjection(Frame &CurrentFrame, KeyFrame *pKF, sAlreadyFound, th, ORBdist) -----------
    def SearchByProjectionKeyFrame(self, cur: Frame, camera: Camera, Tcw_cur, log_scale_factor: float, kf_valid, kf_xyz,
                                   kf_max_dist, kf_min_dist, kf_max_d, kf_angle, kf_desc, th: float, ORBdist: int) -> int:
        """src/ORBmatcher.cc:3809-3937 over flat arrays; cur.mvpMapPoints updated in place."""
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        sf = f32(cur.mvScaleFactors)
        fmp = np.ascontiguousarray(cur.mvpMapPoints, dtype=np.int32)
        val = np.ascontiguousarray(kf_valid, dtype=np.int32)
        xyz, mx, mn, md, ang = f32(kf_xyz), f32(kf_max_dist), f32(kf_min_dist), f32(kf_max_d), f32(kf_angle)
        desc = np.ascontiguousarray(kf_desc, dtype=np.uint8)
        tc = f32(Tcw_cur)
        nm = C.c_int(0)
        check_m(self._h, lib.orbm_search_by_projection_keyframe_host(
            self._h, cur.mvKeysUn.ctypes.data, cur.mDescriptors.ctypes.data, cur.N, cur.bounds, sf.ctypes.data, len(sf),
            float(log_scale_factor), camera, tc.ctypes.data, val.ctypes.data, xyz.ctypes.data, mx.ctypes.data,
            mn.ctypes.data, md.ctypes.data, ang.ctypes.data, desc.ctypes.data, len(val), float(th), int(ORBdist),
            int(self.mbCheckOrientation), fmp.ctypes.data, C.byref(nm)))
        cur.mvpMapPoints = fmp
        return nm.value

    # -- SearchByProjection(KeyFrame*, Scw, vpPoints, vLoopMPCams, vpMatched, th, CalibMatrix) -------
    def SearchByProjectionSim3(self, kf: Frame, kf_cam, camera: Camera, log_scale_factor: float, Scw, calib, mp_valid, mp_xyz,
                               mp_normal, mp_max_dist, mp_min_dist, mp_max_d, mp_desc, th: int) -> int:
        """src/ORBmatcher.cc:566-752 over flat arrays; kf.mvpMapPoints plays vpMatched (updated in place)."""
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        sf = f32(kf.mvScaleFactors)
        cam_of = np.ascontiguousarray(kf_cam, dtype=np.int32)
        matched = np.ascontiguousarray(kf.mvpMapPoints, dtype=np.int32)
        val = np.ascontiguousarray(mp_valid, dtype=np.int32)
        xyz, nrm, mx, mn, md = f32(mp_xyz), f32(mp_normal), f32(mp_max_dist), f32(mp_min_dist), f32(mp_max_d)
        desc = np.ascontiguousarray(mp_desc, dtype=np.uint8)
        S, cal = f32(Scw), f32(calib)
        nm = C.c_int(0)
        check_m(self._h, lib.orbm_search_by_projection_sim3_host(
            self._h, kf.mvKeysUn.ctypes.data, kf.mDescriptors.ctypes.data, cam_of.ctypes.data, kf.N, kf.bounds, sf.ctypes.data,
            len(sf), float(log_scale_factor), camera, S.ctypes.data, cal.ctypes.data, val.ctypes.data, xyz.ctypes.data,
            nrm.ctypes.data, mx.ctypes.data, mn.ctypes.data, md.ctypes.data, desc.ctypes.data, len(val), int(th),
            matched.ctypes.data, C.byref(nm)))
        kf.mvpMapPoints = matched
        return nm.value

    # -- Fuse(KeyFrame*, vpMapPoints, CalibMatrix, th) (src/ORBmatcher.cc:1986-2190), the search part ----------
    def Fuse(self, kf: Frame, kf_cam, camera: Camera, log_scale_factor: float, inv_level_sigma2, Tcw, Ow, calib, mp_valid,
             mp_xyz, mp_normal, mp_max_dist, mp_min_dist, mp_max_d, mp_desc, th: float = 3.0):
        """kf: the key frame (mvKeysUn_total, descriptors, mvuRight_total, mvScaleFactors); Ow = camera centres of
        both cameras [2, 3].  Returns (nFused, best_idx [n_mp, 2]): the key-frame feature each map point fuses
        with per camera (-1 = none); the caller applies Replace / AddObservation in order (:2160-2186)."""
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        sf, isg = f32(kf.mvScaleFactors), f32(inv_level_sigma2)
        if len(sf) != len(isg):
            raise ValueError("mvScaleFactors and mvInvLevelSigma2 must have nlevels entries each")
        cam_of = None if kf_cam is None else np.ascontiguousarray(kf_cam, dtype=np.int32)
        ur = f32(kf.mvuRight if kf.mvuRight is not None else np.full(kf.N, -1.0))
        val = np.ascontiguousarray(mp_valid, dtype=np.int32)
        xyz, nrm, mx, mn, md = f32(mp_xyz), f32(mp_normal), f32(mp_max_dist), f32(mp_min_dist), f32(mp_max_d)
        desc = np.ascontiguousarray(mp_desc, dtype=np.uint8)
        T, O, cal = f32(Tcw), f32(Ow).reshape(6), f32(calib)
        best = np.empty((len(val), 2), dtype=np.int32)
        nf = C.c_int(0)
        check_m(self._h, lib.orbm_fuse_host(
            self._h, kf.mvKeysUn.ctypes.data, kf.mDescriptors.ctypes.data, ur.ctypes.data,
            None if cam_of is None else cam_of.ctypes.data, kf.N, kf.bounds, sf.ctypes.data, isg.ctypes.data, len(sf),
            float(log_scale_factor), camera, T.ctypes.data, O.ctypes.data, cal.ctypes.data, val.ctypes.data, xyz.ctypes.data,
            nrm.ctypes.data, mx.ctypes.data, mn.ctypes.data, md.ctypes.data, desc.ctypes.data, len(val), float(th),
            best.ctypes.data, C.byref(nf)))
        return nf.value, best

    def FuseSim3(self, kf: Frame, kf_cam, camera: Camera, log_scale_factor: float, Scw, calib, mp_valid, mp_xyz, mp_normal,
                 mp_max_dist, mp_min_dist, mp_max_d, mp_desc, th: float = 4.0):
        """Fuse(KeyFrame*, Scw, vpPoints, ..., CalibMatrix) (src/ORBmatcher.cc:2211-2441), the search part.
        Returns (nFused, best_idx [n_mp, 2])."""
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        sf = f32(kf.mvScaleFactors)
        cam_of = None if kf_cam is None else np.ascontiguousarray(kf_cam, dtype=np.int32)
        val = np.ascontiguousarray(mp_valid, dtype=np.int32)
        xyz, nrm, mx, mn, md = f32(mp_xyz), f32(mp_normal), f32(mp_max_dist), f32(mp_min_dist), f32(mp_max_d)
        desc = np.ascontiguousarray(mp_desc, dtype=np.uint8)
        S, cal = f32(Scw), f32(calib)
        best = np.empty((len(val), 2), dtype=np.int32)
        nf = C.c_int(0)
        check_m(self._h, lib.orbm_fuse_sim3_host(
            self._h, kf.mvKeysUn.ctypes.data, kf.mDescriptors.ctypes.data, None if cam_of is None else cam_of.ctypes.data, kf.N,
            kf.bounds, sf.ctypes.data, len(sf), float(log_scale_factor), camera, S.ctypes.data, cal.ctypes.data, val.ctypes.data,
            xyz.ctypes.data, nrm.ctypes.data, mx.ctypes.data, mn.ctypes.data, md.ctypes.data, desc.ctypes.data, len(val), float(th),
            best.ctypes.data, C.byref(nf)))
        return nf.value, best

    # -- DBoW2 vocabulary transform (Frame::ComputeBoW, src/Frame.cc:649-659) ----------------------------------
    def set_vocabulary(self, child_start, child_ids, node_desc, word_id, node_weight, L: int) -> None:
        """Upload the vocabulary tree (see include/orb_b200.h: orbm_set_vocabulary; synth.random_vocabulary)."""
        cs, ci = np.ascontiguousarray(child_start, dtype=np.int32), np.ascontiguousarray(child_ids, dtype=np.int32)
        nd = np.ascontiguousarray(node_desc, dtype=np.uint8).reshape(-1, 32)
        wi, ww = np.ascontiguousarray(word_id, dtype=np.int32), np.ascontiguousarray(node_weight, dtype=np.float64)
        if len(cs) != len(nd) + 1 or len(wi) != len(nd) or len(ww) != len(nd):
            raise ValueError("vocabulary arrays must have n_nodes (+1 for child_start) entries")
        check_m(self._h, lib.orbm_set_vocabulary(self._h, cs.ctypes.data, ci.ctypes.data, nd.ctypes.data, wi.ctypes.data,
                                                 ww.ctypes.data, len(nd), int(L)))

    def ComputeBoW(self, desc, levelsup: int = 4):
        """mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, levelsup).  Returns a dict: word, node,
        weight (per feature), bow = (word ids ascending, L1-normalised values), featvec = (node ids, start, items)."""
        d = np.ascontiguousarray(desc, dtype=np.uint8).reshape(-1, 32)
        n = len(d)
        word, node, weight = np.empty(n, np.int32), np.empty(n, np.int32), np.empty(n, np.float64)
        bw, bv = np.empty(max(n, 1), np.int32), np.empty(max(n, 1), np.float64)
        fn, fs, fi = np.empty(max(n, 1), np.int32), np.empty(n + 1, np.int32), np.empty(max(n, 1), np.int32)
        nb, nf = C.c_int(0), C.c_int(0)
        check_m(self._h, lib.orbm_bow_transform_host(self._h, d.ctypes.data, n, int(levelsup), word.ctypes.data, node.ctypes.data,
                                                     weight.ctypes.data, bw.ctypes.data, bv.ctypes.data, C.byref(nb),
                                                     fn.ctypes.data, fs.ctypes.data, fi.ctypes.data, C.byref(nf)))
        return dict(word=word, node=node, weight=weight, bow=(bw[: nb.value].copy(), bv[: nb.value].copy()),
                    featvec=(fn[: nf.value].copy(), fs[: nf.value + 1].copy(), fi[: fs[nf.value]].copy()))

    # -- SearchBySim3 (src/ORBmatcher.cc:2814-3136) -------------------------------------------------------------
    def SearchBySim3(self, kf1: Frame, cam1, T1w, kf2: Frame, cam2, T2w, camera: Camera, log_scale_factor: float, s12: float,
                     R12, t12, calib, mp1, mp2, th: float = 7.5):
        """mpX = dict(valid, xyz, max_dist, min_dist, max_d, desc) aligned with the keypoints of key frame X.
        Returns (nFound, match12 [n1]) with match12[i1] = key-frame-2 feature of each new mutual match or -1."""
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        i32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.int32)
        sf = f32(kf1.mvScaleFactors)
        c1, c2 = i32(cam1), i32(cam2)
        A = [f32(T1w), f32(T2w), f32(R12), f32(t12), f32(calib)]
        P = []
        for mp in (mp1, mp2):
            P += [i32(mp["valid"]), f32(mp["xyz"]), f32(mp["max_dist"]), f32(mp["min_dist"]), f32(mp["max_d"]),
                  np.ascontiguousarray(mp["desc"], dtype=np.uint8)]
        m12 = np.empty(kf1.N, dtype=np.int32)
        nf = C.c_int(0)
        p = lambda a: None if a is None else a.ctypes.data
        check_m(self._h, lib.orbm_search_by_sim3_host(
            self._h, p(kf1.mvKeysUn), p(kf1.mDescriptors), p(c1), kf1.N, p(A[0]), p(kf2.mvKeysUn), p(kf2.mDescriptors), p(c2),
            kf2.N, p(A[1]), kf1.bounds, p(sf), len(sf), float(log_scale_factor), camera, float(s12), p(A[2]), p(A[3]), p(A[4]),
            *[p(a) for a in P], float(th), p(m12), C.byref(nf)))
        return nf.value, m12

    # -- SearchByBoW, Frame and KeyFrame variants (src/ORBmatcher.cc:206-388, 390-565, 996-1163, 1180-1363) --
    def SearchByBoW(self, desc1, angle1, valid1, featvec1, desc2, angle2, valid2, featvec2, *, keyframe_pair: bool = False):
        """Side 1 = key frame whose map points are searched (valid1[i] = map point exists and is not
        bad), side 2 = frame (keyframe_pair=False: `bestDist1<=TH_LOW`, :324) or second key frame
        (keyframe_pair=True: `bestDist1<TH_LOW`, :1107).  featvec = (node_ids ascending, start, items),
        the CSR form of DBoW2::FeatureVector (see synth.feature_vector).  For the _cam1 variants clear
        valid1/valid2 for indices >= N.  Returns (nmatches, matches12 [n1], matches21 [n2])."""
        d1 = np.ascontiguousarray(desc1, dtype=np.uint8).reshape(-1, 32)
        d2 = np.ascontiguousarray(desc2, dtype=np.uint8).reshape(-1, 32)
        a1 = np.ascontiguousarray(angle1, dtype=np.float32)
        a2 = np.ascontiguousarray(angle2, dtype=np.float32)
        n1, n2 = len(d1), len(d2)
        if len(a1) != n1 or len(a2) != n2:
            raise ValueError("angle arrays must match the descriptor counts")
        v1 = None if valid1 is None else np.ascontiguousarray(valid1, dtype=np.int32)
        v2 = None if valid2 is None else np.ascontiguousarray(valid2, dtype=np.int32)
        keep = []

        def fv(t):
            arrs = [np.ascontiguousarray(x, dtype=np.int32) for x in t]
            if len(arrs[1]) != len(arrs[0]) + 1:
                raise ValueError("feature vector: start must have n_nodes + 1 entries")
            keep.extend(arrs)
            return _lib.FeatVec(arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, len(arrs[0]))

        m12 = np.empty(n1, dtype=np.int32)
        m21 = np.empty(n2, dtype=np.int32)
        nm = C.c_int(0)
        check_m(self._h, lib.orbm_search_by_bow_host(
            self._h, d1.ctypes.data, a1.ctypes.data, None if v1 is None else v1.ctypes.data, n1, fv(featvec1),
            d2.ctypes.data, a2.ctypes.data, None if v2 is None else v2.ctypes.data, n2, fv(featvec2),
            self.mfNNratio, int(self.mbCheckOrientation), self.TH_LOW - 1 if keyframe_pair else self.TH_LOW,
            m12.ctypes.data, m21.ctypes.data, C.byref(nm)))
        return nm.value, m12, m21

    def SearchByBoW_batch(self, pairs, *, keyframe_pair: bool = False):
        """pairs: sequence of (desc1, angle1, valid1, featvec1, desc2, angle2, valid2, featvec2) tuples, one per
        independent (key frame, frame) pair; one library call, the ordered searches run concurrently on the
        GPU.  Returns a list of (nmatches, matches12, matches21)."""
        keep, structs, outs = [], (_lib.BowPair * len(pairs))(), []
        c = lambda a, t: np.ascontiguousarray(a, dtype=t)

        def fv(t):
            arrs = [c(x, np.int32) for x in t]
            if len(arrs[1]) != len(arrs[0]) + 1:
                raise ValueError("feature vector: start must have n_nodes + 1 entries")
            keep.extend(arrs)
            return _lib.FeatVec(arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, len(arrs[0]))

        for i, (d1, a1, v1, f1, d2, a2, v2, f2) in enumerate(pairs):
            d1, d2 = c(d1, np.uint8).reshape(-1, 32), c(d2, np.uint8).reshape(-1, 32)
            a1, a2 = c(a1, np.float32), c(a2, np.float32)
            if len(a1) != len(d1) or len(a2) != len(d2):
                raise ValueError("angle arrays must match the descriptor counts")
            v1 = None if v1 is None else c(v1, np.int32)
            v2 = None if v2 is None else c(v2, np.int32)
            m12, m21 = np.empty(len(d1), np.int32), np.empty(len(d2), np.int32)
            keep.extend([d1, d2, a1, a2, v1, v2])
            outs.append((m12, m21))
            structs[i] = _lib.BowPair(d1.ctypes.data, a1.ctypes.data, None if v1 is None else v1.ctypes.data, len(d1), fv(f1),
                                      d2.ctypes.data, a2.ctypes.data, None if v2 is None else v2.ctypes.data, len(d2), fv(f2),
                                      m12.ctypes.data, m21.ctypes.data, 0)
        check_m(self._h, lib.orbm_search_by_bow_batch_host(self._h, structs, len(pairs), self.mfNNratio,
                                                           int(self.mbCheckOrientation),
                                                           self.TH_LOW - 1 if keyframe_pair else self.TH_LOW))
        return [(structs[i].nmatches, outs[i][0], outs[i][1]) for i in range(len(pairs))]

    # -- SearchForTriangulation (src/ORBmatcher.cc:1364-1720) ---------------------------------------------------
    def SearchForTriangulation(self, k1, desc1, has_mp1, cam1, uright1, featvec1, k2, desc2, has_mp2, cam2, uright2, featvec2,
                               F12s, epipoles, scale_factors2, level_sigma2_2, bOnlyStereo: bool = False, vbCam=(True, True)):
        """Key frames as flat arrays (see include/orb_b200.h: orbm_search_for_triangulation_host); F12s [2,3,3]
        and epipoles [4] come from the caller's pose algebra.  Returns (nmatches, vMatches12 [n1],
        vMatchedPairs [nmatches, 2])."""
        c = lambda a, t: np.ascontiguousarray(a, dtype=t)
        k1, k2 = c(k1, KP_DTYPE), c(k2, KP_DTYPE)
        d1, d2 = c(desc1, np.uint8).reshape(-1, 32), c(desc2, np.uint8).reshape(-1, 32)
        n1, n2 = len(k1), len(k2)
        side1 = [c(has_mp1, np.int32), c(cam1, np.int32), c(uright1, np.float32)]
        side2 = [c(has_mp2, np.int32), c(cam2, np.int32), c(uright2, np.float32)]
        if any(len(a) != n1 for a in side1 + [d1]) or any(len(a) != n2 for a in side2 + [d2]):
            raise ValueError("per-keypoint arrays must match the keypoint counts")
        F = c(F12s, np.float32).reshape(2, 3, 3)
        epi = c(epipoles, np.float32).reshape(4)
        sf, ls = c(scale_factors2, np.float32), c(level_sigma2_2, np.float32)
        if len(sf) != len(ls):
            raise ValueError("scale_factors2 and level_sigma2_2 must have nlevels entries each")
        en = c([int(bool(v)) for v in vbCam], np.int32)
        keep = []

        def fv(t):
            arrs = [np.ascontiguousarray(x, dtype=np.int32) for x in t]
            if len(arrs[1]) != len(arrs[0]) + 1:
                raise ValueError("feature vector: start must have n_nodes + 1 entries")
            keep.extend(arrs)
            return _lib.FeatVec(arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, len(arrs[0]))

        m12 = np.empty(n1, dtype=np.int32)
        nm = C.c_int(0)
        p = lambda a: a.ctypes.data
        check_m(self._h, lib.orbm_search_for_triangulation_host(
            self._h, p(k1), p(d1), p(side1[0]), p(side1[1]), p(side1[2]), n1, fv(featvec1),
            p(k2), p(d2), p(side2[0]), p(side2[1]), p(side2[2]), n2, fv(featvec2), p(F), p(epi), p(sf), p(ls), len(sf),
            int(bOnlyStereo), p(en), int(self.mbCheckOrientation), p(m12), C.byref(nm)))
        idx = np.nonzero(m12 >= 0)[0]
        return nm.value, m12, np.stack([idx, m12[idx]], axis=1)


    # -- MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:325-438) --------------------------------------
    def ComputeDistinctiveDescriptors(self, desc, offsets):
        """Batched over map points: rows offsets[p]:offsets[p+1] of desc are the observed descriptors of
        point p.  Returns best_idx [P] (relative to offsets[p]; -1 for an empty set)."""
        d = np.ascontiguousarray(desc, dtype=np.uint8).reshape(-1, 32)
        off = np.ascontiguousarray(offsets, dtype=np.int32)
        if len(off) < 1 or off[0] != 0 or off[-1] != len(d) or (np.diff(off) < 0).any():
            raise ValueError("offsets must start at 0, be non-decreasing and end at the descriptor count")
        best = np.empty(len(off) - 1, dtype=np.int32)
        check_m(self._h, lib.orbm_compute_distinctive_descriptors_host(self._h, d.ctypes.data, off.ctypes.data, len(off) - 1,
                                                                       best.ctypes.data))
        return best

    def SearchForTriangulation_batch(self, scenes, bOnlyStereo: bool = False, vbCam=(True, True)):
        """scenes: sequence of dicts with the keys of synth.triangulation_scene plus "fv1"/"fv2" (CSR feature
        vectors): one independent key-frame pair each, all with the same number of pyramid levels.  One
        library call.  Returns a list of (nmatches, vMatches12)."""
        c = lambda a, t: np.ascontiguousarray(a, dtype=t)
        keep, structs, outs = [], (_lib.TriPair * len(scenes))(), []
        nlevels = None

        def fv(t):
            arrs = [c(x, np.int32) for x in t]
            keep.extend(arrs)
            return _lib.FeatVec(arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, len(arrs[0]))

        p = lambda a: a.ctypes.data
        for i, sc in enumerate(scenes):
            k1, k2 = c(sc["k1"], KP_DTYPE), c(sc["k2"], KP_DTYPE)
            a = [k1, c(sc["d1"], np.uint8), c(sc["has_mp1"], np.int32), c(sc["cam1"], np.int32), c(sc["uright1"], np.float32),
                 k2, c(sc["d2"], np.uint8), c(sc["has_mp2"], np.int32), c(sc["cam2"], np.int32), c(sc["uright2"], np.float32),
                 c(sc["F12s"], np.float32), c(sc["epipoles"], np.float32), c(sc["scale_factors"], np.float32),
                 c(sc["level_sigma2"], np.float32)]
            if nlevels is None:
                nlevels = len(a[12])
            if len(a[12]) != nlevels or len(a[13]) != nlevels:
                raise ValueError("all pairs must share the number of pyramid levels")
            m12 = np.empty(len(k1), np.int32)
            keep.extend(a)
            outs.append(m12)
            structs[i] = _lib.TriPair(p(a[0]), p(a[1]), p(a[2]), p(a[3]), p(a[4]), len(k1), fv(sc["fv1"]),
                                      p(a[5]), p(a[6]), p(a[7]), p(a[8]), p(a[9]), len(k2), fv(sc["fv2"]),
                                      p(a[10]), p(a[11]), p(a[12]), p(a[13]), p(m12), 0)
        if not scenes:
            return []
        en = c([int(bool(v)) for v in vbCam], np.int32)
        check_m(self._h, lib.orbm_search_for_triangulation_batch_host(self._h, structs, len(scenes), nlevels, int(bOnlyStereo),
                                                                      p(en), int(self.mbCheckOrientation)))
        return [(structs[i].nmatches, outs[i]) for i in range(len(scenes))]
