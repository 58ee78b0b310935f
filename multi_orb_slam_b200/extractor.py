"""Host-side mirror of the reference's `ORB_SLAM2::ORBextractor` (include/ORBextractor.h:45-112,
src/ORBextractor.cc) over the C-ABI CUDA library.  Same constructor arguments, call operator and
getters; batches of frames are the B200-native addition (one launch group covers all frames)."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, Config, check_x, lib, ptr


class ORBextractor:
    HARRIS_SCORE, FAST_SCORE = 0, 1  # enum of include/ORBextractor.h:49

    def __init__(self, nfeatures: int, scaleFactor: float, nlevels: int, iniThFAST: int, minThFAST: int, *,
                 image_size: Optional[Tuple[int, int]] = None, max_batch: int = 1, device: int = -1):
        """image_size=(width, height) sizes the device workspace up front; otherwise it is created
        on the first call from that image's size (like the reference, which allocates per call)."""
        self.nfeatures, self.scaleFactor, self.nlevels = int(nfeatures), float(scaleFactor), int(nlevels)
        self.iniThFAST, self.minThFAST = int(iniThFAST), int(minThFAST)
        self.max_batch, self.device = int(max_batch), int(device)
        self._h = None
        self._size = None
        if image_size is not None:
            self._create(int(image_size[0]), int(image_size[1]))

    # -- handle ---------------------------------------------------------------------------------
    def _create(self, width: int, height: int) -> None:
        self.close()
        cfg = Config(self.nfeatures, self.scaleFactor, self.nlevels, self.iniThFAST, self.minThFAST, width, height,
                     self.max_batch, self.device)
        h = C.c_void_p()
        rc = lib.orbx_create(C.byref(cfg), C.byref(h))
        if rc != _lib.OK:
            raise _lib.OrbError(rc, (lib.orbx_last_error(None) or b"").decode())
        self._h, self._size = h, (width, height)

    def _ensure(self, width: int, height: int) -> None:
        if self._h is None or self._size != (width, height):
            self._create(width, height)

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.orbx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def capacity(self) -> int:
        """Per-frame output capacity: nfeatures + 3*nlevels (octree overshoot, ORBextractor.cc:731)."""
        return self.nfeatures + 3 * self.nlevels

    # -- reference API --------------------------------------------------------------------------
    def __call__(self, image: np.ndarray, mask=None):
        """operator()(image, mask, keypoints, descriptors) — ORBextractor.cc:1044-1107.  The mask is
        ignored, as in the reference.  Returns (keypoints[KP_DTYPE], descriptors[N,32] u8)."""
        if image is None or image.size == 0:
            return np.zeros(0, dtype=KP_DTYPE), np.zeros((0, 32), dtype=np.uint8)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise ValueError("image must be a 2-D uint8 array (CV_8UC1)")  # assert at :1051
        if image.strides[1] != 1:
            image = np.ascontiguousarray(image)
        rows, cols = image.shape
        self._ensure(cols, rows)
        cap = self.capacity
        kps = np.empty(cap, dtype=KP_DTYPE)
        desc = np.empty((cap, 32), dtype=np.uint8)
        n = C.c_int(0)
        check_x(self._h, lib.orbx_extract(self._h, image.ctypes.data, rows, cols, image.strides[0], kps.ctypes.data,
                                          desc.ctypes.data, cap, C.byref(n)))
        return kps[: n.value].copy(), desc[: n.value].copy()

    def GetLevels(self) -> int:
        return self.nlevels

    def GetScaleFactor(self) -> float:
        return self.scaleFactor

    def _tables(self):
        if self._h is None:
            raise RuntimeError("scale tables need a created handle (pass image_size= or extract once)")
        out = [np.zeros(self.nlevels, dtype=np.float32) for _ in range(4)]
        check_x(self._h, lib.orbx_get_scale_tables(self._h, *[o.ctypes.data for o in out]))
        return out

    def GetScaleFactors(self):
        return self._tables()[0]

    def GetInverseScaleFactors(self):
        return self._tables()[1]

    def GetScaleSigmaSquares(self):
        return self._tables()[2]

    def GetInverseScaleSigmaSquares(self):
        return self._tables()[3]

    def features_per_level(self) -> np.ndarray:
        out = np.zeros(self.nlevels, dtype=np.int32)
        check_x(self._h, lib.orbx_get_features_per_level(self._h, out.ctypes.data))
        return out

    def pyramid_view(self):
        """Device view (orbx_pyramid_view) of mvImagePyramid for the last batch; valid until the next extraction."""
        from ._lib import PyramidView
        v = PyramidView()
        check_x(self._h, lib.orbx_get_pyramid_view(self._h, C.byref(v)))
        return v

    def pyramid_level(self, level: int, frame: int = 0, with_border: bool = False) -> np.ndarray:
        """mvImagePyramid[level] of a frame of the last call (public member, ORBextractor.h:92)."""
        w, h = C.c_int(), C.c_int()
        check_x(self._h, lib.orbx_get_pyramid_level(self._h, frame, level, int(with_border), None, 0, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), dtype=np.uint8)
        check_x(self._h, lib.orbx_get_pyramid_level(self._h, frame, level, int(with_border), out.ctypes.data,
                                                    out.strides[0], C.byref(w), C.byref(h)))
        return out

    @property
    def mvImagePyramid(self):
        return [self.pyramid_level(l) for l in range(self.nlevels)]

    # -- batched API ----------------------------------------------------------------------------
    def extract_batch(self, images: np.ndarray):
        """images: [F, H, W] uint8 in host memory (pinned memory makes the copies asynchronous).
        Returns (kps [F, cap] KP_DTYPE, desc [F, cap, 32] u8, counts [F] i32)."""
        if images.dtype != np.uint8 or images.ndim != 3:
            raise ValueError("images must be [F, H, W] uint8")
        if images.strides[2] != 1:
            images = np.ascontiguousarray(images)
        F, rows, cols = images.shape
        self._ensure(cols, rows)
        cap = self.capacity
        kps = np.empty((F, cap), dtype=KP_DTYPE)
        desc = np.empty((F, cap, 32), dtype=np.uint8)
        counts = np.zeros(F, dtype=np.int32)
        check_x(self._h, lib.orbx_extract_batch_host(self._h, images.ctypes.data, F, images.strides[0], images.strides[1],
                                                     kps.ctypes.data, desc.ctypes.data, counts.ctypes.data, cap))
        return kps, desc, counts

    def extract_batch_device(self, images, kps=None, desc=None, counts=None):
        """images: torch uint8 CUDA tensor [F, H, W] (row-contiguous).  Outputs are torch CUDA tensors
        (allocated here unless given): kps [F, cap, 6] f32 (column 5 holds the octave's int32 bits),
        desc [F, cap, 32] u8, counts [F] i32.  Asynchronous on the handle's stream (see `stream`);
        the caller's current stream must have finished producing `images`."""
        import torch
        assert images.is_cuda and images.dtype == torch.uint8 and images.dim() == 3 and images.stride(2) == 1
        F, rows, cols = images.shape
        self._ensure(cols, rows)
        cap = self.capacity
        if kps is None:
            kps = torch.empty((F, cap, 6), dtype=torch.float32, device=images.device)
        if desc is None:
            desc = torch.empty((F, cap, 32), dtype=torch.uint8, device=images.device)
        if counts is None:
            counts = torch.empty((F,), dtype=torch.int32, device=images.device)
        check_x(self._h, lib.orbx_extract_batch_device(self._h, images.data_ptr(), F, images.stride(0), images.stride(1),
                                                       kps.data_ptr(), desc.data_ptr(), counts.data_ptr(), cap))
        # The kernels run asynchronously on the handle's stream and level 0 of the pyramid (pyramid_view,
        # pyramid_level) IS this buffer: hold it until the next batch so a temporary cannot be recycled under them.
        self._last_images = images
        return kps, desc, counts

    def sync(self) -> None:
        check_x(self._h, lib.orbx_sync(self._h))

    @property
    def stream(self) -> int:
        """cudaStream_t of the handle (wrap with torch.cuda.ExternalStream to record events on it)."""
        return lib.orbx_stream(self._h)

    def set_stream(self, cuda_stream: int) -> None:
        """Run on a caller-owned cudaStream_t (e.g. torch.cuda.Stream().cuda_stream); 0/None restores."""
        check_x(self._h, lib.orbx_set_stream(self._h, cuda_stream or None))

    @property
    def launch_count(self) -> int:
        return lib.orbx_launch_count(self._h) if self._h else 0

    def set_profiling(self, on: bool) -> None:
        check_x(self._h, lib.orbx_set_profiling(self._h, int(on)))

    STAGES = ("pyramid", "fast", "octree", "blur", "orient_describe")

    def stage_times_ms(self):
        """(summed ms per stage since set_profiling(True), number of launch groups)."""
        out = np.zeros(5, dtype=np.float64)
        n = C.c_longlong(0)
        check_x(self._h, lib.orbx_stage_times_ms(self._h, out.ctypes.data, C.byref(n)))
        return out, n.value

    # -- stage taps (parity tests) ----------------------------------------------------------------
    def debug_candidates(self, level: int, frame: int = 0):
        cap = 1 << 18
        x, y, s = (np.zeros(cap, dtype=np.int32) for _ in range(3))
        n = C.c_int()
        check_x(self._h, lib.orbx_debug_candidates(self._h, frame, level, x.ctypes.data, y.ctypes.data, s.ctypes.data, cap,
                                                   C.byref(n)))
        return x[: n.value].copy(), y[: n.value].copy(), s[: n.value].copy()

    def debug_blurred(self, level: int, frame: int = 0) -> np.ndarray:
        lv = self.pyramid_level(level, frame)
        out = np.empty_like(lv)
        check_x(self._h, lib.orbx_debug_blurred(self._h, frame, level, out.ctypes.data, out.strides[0]))
        return out
