"""Streaming front end of a multi-camera rig: pinned host frames in, pinned host features out.

Thin Python view of the C-ABI pipeline (`orbp_*`, include/orb_b200.h; multi_orb_slam_b200/csrc/pipeline_api.cu) —
the same object a C++ Tracking would drive.  What Tracking does per rig-frame in the reference (Frame::Frame runs one
ORBextractor per camera, src/Frame.cc:148-346; MonocularInitialization matches consecutive frames of camera 1 with
ORBmatcher::SearchForInitialization, src/Tracking.cc:870-871), batched over `rig_frames` rig-frames per step and
software-pipelined over four CUDA streams inside the library (copy-in, extraction, matching on a high-priority side
stream, copy-out; `depth` steps in flight).  `submit` never blocks the host, `result` waits for one step.
There is no CPU fallback: the extractor and matcher are the CUDA library's."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, PipelineConfig, PipelineResult, lib


@dataclass
class RigStepResult:
    """Pinned host arrays of one step, owned by the pipeline (valid until the slot is reused `depth` submits later)."""
    kps: list        # per camera [F, cap, 6] f32 torch view (orbx_keypoint rows; column 5 = octave bits)
    desc: list       # per camera [F, cap, 32] u8
    counts: list     # per camera [F] i32
    matches12: object  # [F-1, cap0] i32: SearchForInitialization(frame t, frame t+1) of camera 0 (None without match)
    nmatches: object   # [F-1] i32

    def keypoints(self, cam: int, frame: int) -> np.ndarray:
        """KP_DTYPE view of one camera-frame's valid keypoints."""
        n = int(self.counts[cam][frame])
        return self.kps[cam][frame, :n].numpy().view(KP_DTYPE).reshape(-1)


class RigPipeline:
    def __init__(self, nfeatures: Sequence[int] = (1000, 500), scaleFactor: float = 1.2, nlevels: int = 8,
                 iniThFAST: int = 20, minThFAST: int = 7, *, image_size: Tuple[int, int] = (640, 480),
                 rig_frames: int = 256, n_chunks: int = 1, depth: int = 3, window: int = 100, nnratio: float = 0.9,
                 match: bool = True, device: int = 0):
        import torch
        self.torch = torch
        if n_chunks != 1:
            raise ValueError("the C-ABI pipeline submits a step in one piece (n_chunks = 1)")
        self.n_chunks = 1
        self.F, self.depth, self.window = int(rig_frames), max(2, int(depth)), int(window)
        self.W, self.H = int(image_size[0]), int(image_size[1])
        self.n_cams = len(nfeatures)
        self.match = bool(match) and self.F > 1
        cfg = PipelineConfig()
        cfg.n_cams = self.n_cams
        for c, nf in enumerate(nfeatures):
            cfg.nfeatures[c] = int(nf)
        cfg.scale_factor, cfg.nlevels, cfg.ini_th_fast, cfg.min_th_fast = float(scaleFactor), int(nlevels), int(iniThFAST), int(minThFAST)
        cfg.width, cfg.height, cfg.rig_frames, cfg.depth = self.W, self.H, self.F, self.depth
        cfg.match, cfg.window, cfg.nnratio, cfg.check_ori, cfg.device = int(self.match), self.window, float(nnratio), 1, int(device)
        h = C.c_void_p()
        rc = lib.orbp_create(C.byref(cfg), C.byref(h))
        if rc != _lib.OK:
            raise _lib.OrbError(rc, (lib.orbp_last_error(None) or b"").decode())
        self._h = h
        self.caps = [lib.orbp_capacity(self._h, c) for c in range(self.n_cams)]
        dev = torch.device("cuda", device)
        self.dev = dev
        # the library's streams, for callers that time or order work against the pipeline
        self.s_in, self.s_compute, self.s_match, self.s_out = (
            torch.cuda.ExternalStream(lib.orbp_stream(self._h, i), device=dev) for i in range(4))
        self.n_submitted = 0
        self._keep = {}
        F = self.F
        self.h2d_bytes_per_step = self.n_cams * F * self.H * self.W
        self.d2h_bytes_per_step = sum(F * c * (24 + 32) + 4 * F for c in self.caps) + \
            ((F - 1) * self.caps[0] * 4 + (F - 1) * 4) * int(self.match)

    def _check(self, rc: int) -> None:
        if rc != _lib.OK:
            raise _lib.OrbError(rc, (lib.orbp_last_error(self._h) or b"").decode())

    def close(self) -> None:
        if getattr(self, "_h", None):
            lib.orbp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return lib.orbp_launch_count(self._h) if self._h else 0

    def submit(self, h_images: Sequence) -> int:
        """h_images: one uint8 tensor / array [F, H, W] per camera in host memory (pinned memory makes the copies
        asynchronous; the same row / frame strides for all cameras).  Returns the step's ticket."""
        assert len(h_images) == self.n_cams
        a0 = h_images[0]
        strides = (a0.stride(0), a0.stride(1)) if hasattr(a0, "stride") else (a0.strides[0], a0.strides[1])
        ptrs = (C.c_void_p * self.n_cams)()
        for c, a in enumerate(h_images):
            assert tuple(a.shape) == (self.F, self.H, self.W)
            st = (a.stride(0), a.stride(1)) if hasattr(a, "stride") else (a.strides[0], a.strides[1])
            assert st == strides, "all cameras must share frame / row strides"
            ptrs[c] = _lib.ptr(a)
        t = lib.orbp_submit(self._h, ptrs, strides[0], strides[1])
        if t < 0:
            self._check(int(t))
        self._keep[t % self.depth] = h_images  # the host frames must outlive their copy-in
        self.n_submitted = t + 1
        return int(t)

    def result(self, ticket: int) -> RigStepResult:
        """Blocks until the step's results are in pinned host memory."""
        torch = self.torch
        r = PipelineResult()
        rc = lib.orbp_wait(self._h, ticket, C.byref(r))
        if rc == _lib.E_STATE:
            raise ValueError("ticket no longer (or not yet) held by the pipeline")
        self._check(rc)
        F = self.F

        def view(ptr, shape, dtype):
            n = int(np.prod(shape))
            ct = {np.float32: C.c_float, np.uint8: C.c_uint8, np.int32: C.c_int32}[dtype]
            arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).reshape(shape)
            return torch.from_numpy(arr)

        kps = [view(r.kps[c], (F, self.caps[c], 6), np.float32) for c in range(self.n_cams)]
        desc = [view(r.desc[c], (F, self.caps[c], 32), np.uint8) for c in range(self.n_cams)]
        counts = [view(r.counts[c], (F,), np.int32) for c in range(self.n_cams)]
        m12 = view(r.matches12, (F - 1, self.caps[0]), np.int32) if self.match else None
        nm = view(r.nmatches, (F - 1,), np.int32) if self.match else None
        return RigStepResult(kps, desc, counts, m12, nm)

    def run(self, h_images: Sequence) -> RigStepResult:
        """One synchronous step (submit + result)."""
        return self.result(self.submit(h_images))

    def drain(self) -> None:
        self._check(lib.orbp_drain(self._h))
