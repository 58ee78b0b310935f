"""ctypes binding of the C-ABI library (include/orb_b200.h).  No CPU fallback: if the CUDA library
has not been built, importing this module raises."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ORB_B200_LIB") or os.path.join(HERE, "liborb_b200.so")  # override: development A/B runs only

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])
MP_DTYPE = np.dtype([("proj_x", "<f4"), ("proj_y", "<f4"), ("proj_xr", "<f4"), ("view_cos", "<f4"),
                     ("level", "<i4"), ("track_in_view", "<i4"), ("bad", "<i4")])

OK, E_INVALID, E_CUDA, E_CAPACITY, E_STATE = 0, -1, -2, -3, -4


class OrbError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"orb_b200 error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32), ("width", C.c_int32),
                ("height", C.c_int32), ("max_batch", C.c_int32), ("device", C.c_int32)]


class Camera(C.Structure):
    _fields_ = [("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("mb", C.c_float),
                ("mbf", C.c_float)]


class FeatVec(C.Structure):
    """orbm_featvec: DBoW2::FeatureVector flattened to CSR (node ids ascending)."""
    _fields_ = [("node_id", C.c_void_p), ("start", C.c_void_p), ("items", C.c_void_p), ("n_nodes", C.c_int32)]


class BowPair(C.Structure):
    """orbm_bow_pair (include/orb_b200.h)."""
    _fields_ = [("desc1", C.c_void_p), ("angle1", C.c_void_p), ("valid1", C.c_void_p), ("n1", C.c_int32), ("fv1", FeatVec),
                ("desc2", C.c_void_p), ("angle2", C.c_void_p), ("valid2", C.c_void_p), ("n2", C.c_int32), ("fv2", FeatVec),
                ("matches12", C.c_void_p), ("matches21", C.c_void_p), ("nmatches", C.c_int32)]


class TriPair(C.Structure):
    """orbm_tri_pair (include/orb_b200.h)."""
    _fields_ = [("k1", C.c_void_p), ("desc1", C.c_void_p), ("has_mp1", C.c_void_p), ("cam1", C.c_void_p), ("uright1", C.c_void_p),
                ("n1", C.c_int32), ("fv1", FeatVec),
                ("k2", C.c_void_p), ("desc2", C.c_void_p), ("has_mp2", C.c_void_p), ("cam2", C.c_void_p), ("uright2", C.c_void_p),
                ("n2", C.c_int32), ("fv2", FeatVec),
                ("F12s", C.c_void_p), ("epipoles", C.c_void_p), ("scale_factors2", C.c_void_p), ("level_sigma2_2", C.c_void_p),
                ("matches12", C.c_void_p), ("nmatches", C.c_int32)]


class Bounds(C.Structure):
    _fields_ = [("min_x", C.c_float), ("max_x", C.c_float), ("min_y", C.c_float), ("max_y", C.c_float)]


class PyramidView(C.Structure):
    """orbx_pyramid_view (include/orb_b200.h): device view of an extractor's mvImagePyramid for its last batch."""
    _fields_ = [("nlevels", C.c_int32), ("n_frames", C.c_int32), ("w", C.c_int32 * 16), ("h", C.c_int32 * 16),
                ("pitch", C.c_int32 * 16), ("base", C.c_void_p * 16), ("frame_stride", C.c_size_t * 16),
                ("scale", C.c_float * 16), ("inv_scale", C.c_float * 16), ("stream", C.c_void_p)]


class PipelineConfig(C.Structure):
    """orbp_config (include/orb_b200.h)."""
    _fields_ = [("n_cams", C.c_int32), ("nfeatures", C.c_int32 * 8), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("ini_th_fast", C.c_int32), ("min_th_fast", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("rig_frames", C.c_int32), ("depth", C.c_int32), ("match", C.c_int32), ("window", C.c_int32),
                ("nnratio", C.c_float), ("check_ori", C.c_int32), ("device", C.c_int32)]


class PipelineResult(C.Structure):
    """orbp_result (include/orb_b200.h)."""
    _fields_ = [("kps", C.c_void_p * 8), ("desc", C.c_void_p * 8), ("counts", C.c_void_p * 8), ("cap", C.c_int32 * 8),
                ("matches12", C.c_void_p), ("nmatches", C.c_void_p), ("rig_frames", C.c_int32)]


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -m multi_orb_slam_b200.build` "
        "(__graft_entry__.build()).  multi_orb_slam_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_SIGS = {
    "orbx_create": (_i, [C.POINTER(Config), C.POINTER(_vp)]),
    "orbx_destroy": (None, [_vp]),
    "orbx_last_error": (C.c_char_p, [_vp]),
    "orbx_get_scale_tables": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "orbx_get_features_per_level": (_i, [_vp, _vp]),
    "orbx_max_keypoints": (_i, [_vp]),
    "orbx_extract": (_i, [_vp, _vp, _i, _i, _sz, _vp, _vp, _i, C.POINTER(_i)]),
    "orbx_extract_batch_host": (_i, [_vp, _vp, _i, _sz, _sz, _vp, _vp, _vp, _i]),
    "orbx_extract_batch_device": (_i, [_vp, _vp, _i, _sz, _sz, _vp, _vp, _vp, _i]),
    "orbx_sync": (_i, [_vp]),
    "orbx_stream": (_vp, [_vp]),
    "orbx_set_stream": (_i, [_vp, _vp]),
    "orbx_get_pyramid_level": (_i, [_vp, _i, _i, _i, _vp, _sz, C.POINTER(_i), C.POINTER(_i)]),
    "orbx_debug_candidates": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, C.POINTER(_i)]),
    "orbx_debug_blurred": (_i, [_vp, _i, _i, _vp, _sz]),
    "orbx_launch_count": (C.c_longlong, [_vp]),
    "orbx_set_profiling": (_i, [_vp, _i]),
    "orbx_stage_times_ms": (_i, [_vp, _vp, C.POINTER(C.c_longlong)]),
    "orbm_create": (_i, [_i, C.POINTER(_vp)]),
    "orbm_destroy": (None, [_vp]),
    "orbm_last_error": (C.c_char_p, [_vp]),
    "orbm_sync": (_i, [_vp]),
    "orbm_set_stream": (_i, [_vp, _vp]),
    "orbm_launch_count": (C.c_longlong, [_vp]),
    "orbm_distance_pairs_host": (_i, [_vp, _vp, _vp, _i, _vp]),
    "orbm_bruteforce_device": (_i, [_vp, _vp, _i, _vp, _i, _f, _i, _vp, _vp, _vp]),
    "orbm_bruteforce_host": (_i, [_vp, _vp, _i, _vp, _i, _f, _i, _vp, _vp, _vp]),
    "orbm_bruteforce_batch_device": (_i, [_vp, _i, _i, _vp, _vp, _sz, _vp, _vp, _sz, _f, _i, _vp, _vp, _vp]),
    "orbm_bruteforce_indexed_device": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _f, _i, _vp, _vp, _vp]),
    "orbm_search_for_initialization_device": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, Bounds, _vp, _i, _f, _i,
                                                  _vp, _vp]),
    "orbm_search_for_initialization_host": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, Bounds, _vp, _i, _f, _i,
                                                _vp, _vp]),
    "orbm_debug_set_row_budget": (_i, [_vp, _i]),
    "orbm_search_by_projection_points_host": (_i, [_vp, _vp, _vp, _vp, _i, Bounds, _vp, _i, _vp, _vp, _vp, _i, _f, _f,
                                                  _vp, _vp, C.POINTER(_i)]),
    "orbm_search_by_projection_frame_host": (_i, [_vp, _vp, _vp, _vp, _vp, _i, Bounds, _vp, _i, Camera, _vp, _vp, _vp, _vp,
                                                 _vp, _vp, _vp, _vp, _i, _vp, _f, _i, _i, _vp, _vp, C.POINTER(_i)]),
    "orbm_search_by_projection_keyframe_host": (_i, [_vp, _vp, _vp, _i, Bounds, _vp, _i, _f, Camera, _vp, _vp, _vp, _vp, _vp,
                                                    _vp, _vp, _vp, _i, _f, _i, _i, _vp, C.POINTER(_i)]),
    "orbm_search_by_projection_sim3_host": (_i, [_vp, _vp, _vp, _vp, _i, Bounds, _vp, _i, _f, Camera, _vp, _vp, _vp, _vp, _vp,
                                                _vp, _vp, _vp, _vp, _i, _i, _vp, C.POINTER(_i)]),
    "orbm_search_by_bow_host": (_i, [_vp, _vp, _vp, _vp, _i, FeatVec, _vp, _vp, _vp, _i, FeatVec, _f, _i, _i, _vp, _vp,
                                    C.POINTER(_i)]),
    "orbm_fuse_host": (_i, [_vp, _vp, _vp, _vp, _vp, _i, Bounds, _vp, _vp, _i, _f, Camera, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                           _vp, _vp, _i, _f, _vp, C.POINTER(_i)]),
    "orbm_fuse_sim3_host": (_i, [_vp, _vp, _vp, _vp, _i, Bounds, _vp, _i, _f, Camera, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i,
                                _f, _vp, C.POINTER(_i)]),
    "orbm_search_by_sim3_host": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, Bounds, _vp, _i, _f, Camera, _f, _vp, _vp,
                                     _vp] + [_vp] * 12 + [_f, _vp, C.POINTER(_i)]),
    "orbm_set_vocabulary": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i]),
    "orbm_bow_transform_host": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, C.POINTER(_i), _vp, _vp, _vp, C.POINTER(_i)]),
    "orbm_compute_distinctive_descriptors_host": (_i, [_vp, _vp, _vp, _i, _vp]),
    "orbm_undistort_keypoints_device": (_i, [_vp, _i, _i, _vp, _vp, _f, _f, _f, _f, _vp, _vp]),
    "orbm_undistort_keypoints_host": (_i, [_vp, _vp, _i, _f, _f, _f, _f, _vp, _vp]),
    "orbm_compute_image_bounds_host": (_i, [_vp, _i, _i, _f, _f, _f, _f, _vp, C.POINTER(Bounds)]),
    "orbx_get_pyramid_view": (_i, [_vp, C.POINTER(PyramidView)]),
    "orbm_compute_stereo_matches_device": (_i, [_vp, C.POINTER(PyramidView), C.POINTER(PyramidView), _i, _i, _vp, _vp, _vp, _i,
                                                _vp, _vp, _vp, _f, _f, _vp, _vp]),
    "orbm_compute_stereo_from_rgbd_device": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _sz, _sz, _f, _vp, _vp]),
    "orbm_assign_features_to_grid_device": (_i, [_vp, _i, _i, _vp, _vp, Bounds, _vp, _vp]),
    "orbm_search_by_bow_batch_host": (_i, [_vp, _vp, _i, _f, _i, _i]),
    "orbm_search_for_triangulation_batch_host": (_i, [_vp, _vp, _i, _i, _i, _vp, _i]),
    "orbm_search_for_triangulation_host": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, FeatVec, _vp, _vp, _vp, _vp, _vp, _i,
                                               FeatVec, _vp, _vp, _vp, _vp, _i, _i, _vp, _i, _vp, C.POINTER(_i)]),
    "orbp_create": (_i, [C.POINTER(PipelineConfig), C.POINTER(_vp)]),
    "orbp_destroy": (None, [_vp]),
    "orbp_last_error": (C.c_char_p, [_vp]),
    "orbp_capacity": (_i, [_vp, _i]),
    "orbp_submit": (C.c_longlong, [_vp, _vp, _sz, _sz]),
    "orbp_wait": (_i, [_vp, C.c_longlong, C.POINTER(PipelineResult)]),
    "orbp_drain": (_i, [_vp]),
    "orbp_stream": (_vp, [_vp, _i]),
    "orbp_launch_count": (C.c_longlong, [_vp]),
    "orbd_get_unique_id": (_i, [_vp]),
    "orbd_comm_create": (_i, [_i, _i, _vp, _i, C.POINTER(_vp)]),
    "orbd_comm_destroy": (None, [_vp]),
    "orbd_rank": (_i, [_vp]),
    "orbd_world": (_i, [_vp]),
    "orbd_allgather_inplace": (_i, [_vp, _vp, _sz, _vp]),
    "orbd_nccl_version": (_i, []),
    "orbd_last_error": (C.c_char_p, [_vp]),
}
EXPORTS = tuple(_SIGS)
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here = header/library mismatch
    _fn.restype, _fn.argtypes = _res, _args


def check_x(handle, rc: int) -> None:
    if rc != OK:
        raise OrbError(rc, (lib.orbx_last_error(handle) or b"").decode())


def check_m(handle, rc: int) -> None:
    if rc != OK:
        raise OrbError(rc, (lib.orbm_last_error(handle) or b"").decode())


def check_d(handle, rc: int) -> None:
    if rc != OK:
        raise OrbError(rc, (lib.orbd_last_error(handle) or b"").decode())


def ptr(a) -> int:
    """Raw address of a numpy array or torch tensor (device pointers pass through unchanged)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()
