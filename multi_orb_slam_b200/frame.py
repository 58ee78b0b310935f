"""Device-resident mirror of the glue in the reference's `Frame` constructor between the extractor
and the matchers (src/Frame.cc): UndistortKeyPoints (:673-706), ComputeImageBounds (:743-779),
ComputeStereoFromRGBD (:959-985) and AssignFeaturesToGrid (:348-395), batched over the frames an
`ORBextractor.extract_batch_device` call produced, so keypoints never leave the GPU between
extraction and matching.  CUDA only (C-ABI library); no CPU fallback."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, Bounds, check_m, lib

FRAME_GRID_COLS, FRAME_GRID_ROWS = 64, 48  # include/Frame.h:36-37


class FrameGlue:
    def __init__(self, fx: float, fy: float, cx: float, cy: float, dist_coef: Sequence[float] = (0, 0, 0, 0, 0), *,
                 mbf: float = 0.0, device: int = -1):
        """fx..cy = mK, dist_coef = mDistCoef (k1, k2, p1, p2[, k3]), mbf = baseline * fx."""
        self.fx, self.fy, self.cx, self.cy, self.mbf = float(fx), float(fy), float(cx), float(cy), float(mbf)
        d = list(dist_coef) + [0.0] * (5 - len(dist_coef))
        self.dist = np.asarray(d[:5], dtype=np.float32)
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

    def set_stream(self, cuda_stream: int) -> None:
        check_m(self._h, lib.orbm_set_stream(self._h, cuda_stream or None))

    def sync(self) -> None:
        check_m(self._h, lib.orbm_sync(self._h))

    # -- Frame::ComputeImageBounds ---------------------------------------------------------------------
    def ComputeImageBounds(self, cols: int, rows: int) -> Bounds:
        b = Bounds()
        check_m(self._h, lib.orbm_compute_image_bounds_host(self._h, int(cols), int(rows), self.fx, self.fy, self.cx, self.cy,
                                                            self.dist.ctypes.data, C.byref(b)))
        return b

    # -- Frame::UndistortKeyPoints -----------------------------------------------------------------------
    def UndistortKeyPoints(self, keys: np.ndarray) -> np.ndarray:
        """One frame in host memory: mvKeys (KP_DTYPE) -> mvKeysUn."""
        k = np.ascontiguousarray(keys, dtype=KP_DTYPE)
        out = np.empty_like(k)
        check_m(self._h, lib.orbm_undistort_keypoints_host(self._h, k.ctypes.data, len(k), self.fx, self.fy, self.cx, self.cy,
                                                           self.dist.ctypes.data, out.ctypes.data))
        return out

    def undistort_batch_device(self, kps, counts, out=None):
        """kps [F, cap, 6] f32 CUDA tensor (extractor layout), counts [F] i32 -> kps_un (same layout)."""
        import torch
        F, cap = kps.shape[0], kps.shape[1]
        if out is None:
            out = torch.empty_like(kps)
        check_m(self._h, lib.orbm_undistort_keypoints_device(self._h, F, cap, kps.data_ptr(), counts.data_ptr(), self.fx, self.fy,
                                                             self.cx, self.cy, self.dist.ctypes.data, out.data_ptr()))
        return out

    # -- Frame::ComputeStereoFromRGBD --------------------------------------------------------------------
    def stereo_from_rgbd_batch_device(self, kps, kps_un, counts, depth, uright=None, depth_out=None):
        """depth [F, rows, cols] f32 CUDA tensor (row-contiguous) -> (mvuRight, mvDepth) [F, cap] f32."""
        import torch
        assert depth.is_cuda and depth.dtype == torch.float32 and depth.dim() == 3 and depth.stride(2) == 1
        F, cap = kps.shape[0], kps.shape[1]
        if uright is None:
            uright = torch.empty((F, cap), dtype=torch.float32, device=kps.device)
        if depth_out is None:
            depth_out = torch.empty((F, cap), dtype=torch.float32, device=kps.device)
        check_m(self._h, lib.orbm_compute_stereo_from_rgbd_device(
            self._h, F, cap, kps.data_ptr(), kps_un.data_ptr(), counts.data_ptr(), depth.data_ptr(), depth.shape[2],
            depth.shape[1], depth.stride(1), depth.stride(0), self.mbf, uright.data_ptr(), depth_out.data_ptr()))
        return uright, depth_out

    # -- Frame::ComputeStereoMatches (rectified stereo; src/Frame.cc:782-956, commented upstream code) ---
    def stereo_matches_batch_device(self, ex_left, ex_right, kps_l, desc_l, counts_l, kps_r, desc_r, counts_r, mb=None,
                                    uright=None, depth_out=None):
        """Left/right batch outputs of two ORBextractor handles (whose last batch they are) ->
        (mvuRight, mvDepth) [F, cap_l] f32, all on the device.  mb = baseline in metres (default mbf / fx)."""
        import torch
        F, cap_l, cap_r = kps_l.shape[0], kps_l.shape[1], kps_r.shape[1]
        if uright is None:
            uright = torch.empty((F, cap_l), dtype=torch.float32, device=kps_l.device)
        if depth_out is None:
            depth_out = torch.empty((F, cap_l), dtype=torch.float32, device=kps_l.device)
        vl, vr = ex_left.pyramid_view(), ex_right.pyramid_view()
        check_m(self._h, lib.orbm_compute_stereo_matches_device(
            self._h, C.byref(vl), C.byref(vr), F, cap_l, kps_l.data_ptr(), desc_l.data_ptr(), counts_l.data_ptr(), cap_r,
            kps_r.data_ptr(), desc_r.data_ptr(), counts_r.data_ptr(), self.mbf, float(np.float32(self.mbf) / np.float32(self.fx)) if mb is None else mb,
            uright.data_ptr(), depth_out.data_ptr()))
        return uright, depth_out

    # -- Frame::AssignFeaturesToGrid ---------------------------------------------------------------------
    def assign_features_to_grid_batch_device(self, kps_un, counts, bounds: Bounds, cell_start=None, items=None):
        """-> (cell_start [F, 64*48+1] i32, items [F, cap] i16-as-u16): CSR over cell = ix*48 + iy;
        mGrid[ix][iy] of frame f = items[f, cell_start[f, c]:cell_start[f, c+1]]."""
        import torch
        F, cap = kps_un.shape[0], kps_un.shape[1]
        if cell_start is None:
            cell_start = torch.empty((F, FRAME_GRID_COLS * FRAME_GRID_ROWS + 1), dtype=torch.int32, device=kps_un.device)
        if items is None:
            items = torch.empty((F, cap), dtype=torch.int16, device=kps_un.device)
        check_m(self._h, lib.orbm_assign_features_to_grid_device(self._h, F, cap, kps_un.data_ptr(), counts.data_ptr(), bounds,
                                                                 cell_start.data_ptr(), items.data_ptr()))
        return cell_start, items
