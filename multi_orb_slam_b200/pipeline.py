"""Streaming front end of a multi-camera rig: pinned host frames in, pinned host features out.

What Tracking does per rig-frame in the reference (Frame::Frame runs one ORBextractor per camera,
src/Frame.cc:75-140; MonocularInitialization matches consecutive frames of camera 1 with
ORBmatcher::SearchForInitialization, src/Tracking.cc:870-871), batched over `rig_frames`
rig-frames per step and software-pipelined over three CUDA streams:

    copy-in   H2D of the next chunk of frames            (PCIe, DMA engine)
    compute   extractor of every camera (all extractor kernels on ONE stream: they fill the GPU
              on their own, running cameras concurrently only thrashes the caches)
    match     SearchForInitialization of camera 0, on a high-priority side stream as soon as camera
              0 is extracted: its ordered resolve is a serial chain per pair that leaves the SMs
              ~90 % idle, so it runs underneath the extraction of the other cameras
    copy-out  D2H of keypoints / descriptors / matches    (PCIe, the other DMA engine)

Device image buffers and output buffers are `depth` (default 3) steps deep, so the copies of step
k+1 overlap the kernels of step k even while the consumer still waits for step k-1; `submit` never
blocks the host, `result` waits for one step.
There is no CPU fallback: the extractor and matcher are the CUDA library's."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

from ._lib import Bounds
from .extractor import ORBextractor
from .matcher import ORBmatcher


@dataclass
class RigStepResult:
    """Pinned host tensors of one step (valid until the slot is reused `depth` submits later)."""
    kps: list        # per camera [F, cap, 6] f32 (orbx_keypoint rows; column 5 = octave bits)
    desc: list       # per camera [F, cap, 32] u8
    counts: list     # per camera [F] i32
    matches12: object  # [F-1, cap0] i32: SearchForInitialization(frame t, frame t+1) of camera 0
    nmatches: object   # [F-1] i32


class RigPipeline:
    def __init__(self, nfeatures: Sequence[int] = (1000, 500), scaleFactor: float = 1.2, nlevels: int = 8,
                 iniThFAST: int = 20, minThFAST: int = 7, *, image_size: Tuple[int, int] = (640, 480),
                 rig_frames: int = 256, n_chunks: int = 1, depth: int = 3, window: int = 100, nnratio: float = 0.9,
                 match: bool = True, device: int = 0):
        import torch
        self.torch = torch
        self.F, self.depth, self.window, self.match = int(rig_frames), int(depth), int(window), bool(match)
        self.W, self.H = int(image_size[0]), int(image_size[1])
        self.n_chunks = max(1, min(int(n_chunks), self.F))
        self.chunk = (self.F + self.n_chunks - 1) // self.n_chunks
        self.dev = torch.device("cuda", device)
        self.n_cams = len(nfeatures)
        self.s_in, self.s_compute, self.s_out = (torch.cuda.Stream(device=self.dev) for _ in range(3))
        self.s_match = torch.cuda.Stream(device=self.dev, priority=-1)
        self.ex: List[ORBextractor] = [
            ORBextractor(nf, scaleFactor, nlevels, iniThFAST, minThFAST, image_size=image_size, max_batch=self.chunk,
                         device=device) for nf in nfeatures]
        for e in self.ex:
            e.set_stream(self.s_compute.cuda_stream)
        self.matcher = ORBmatcher(nnratio, True, device=device)
        self.matcher.set_stream(self.s_match.cuda_stream)
        self.caps = [e.capacity for e in self.ex]
        self.bounds = Bounds(0.0, float(self.W), 0.0, float(self.H))
        F, dev = self.F, self.dev
        # ring of chunk-sized device image buffers, `depth` steps deep: with depth 3 the consumer can still be
        # reading step k-2 while step k-1 computes and the frames of step k are already on their way
        self.n_ring = self.depth * self.n_chunks
        self.img = [[torch.empty((self.chunk, self.H, self.W), dtype=torch.uint8, device=dev) for _ in range(self.n_cams)]
                    for _ in range(self.n_ring)]
        self.img_free = [None] * self.n_ring

        def outs(pin):
            kw = dict(device=dev) if not pin else {}
            mk = (lambda *a, **k: torch.empty(*a, **k).pin_memory()) if pin else torch.empty
            return RigStepResult(
                kps=[mk((F, c, 6), dtype=torch.float32, **kw) for c in self.caps],
                desc=[mk((F, c, 32), dtype=torch.uint8, **kw) for c in self.caps],
                counts=[mk((F,), dtype=torch.int32, **kw) for c in self.caps],
                matches12=mk((max(F - 1, 1), self.caps[0]), dtype=torch.int32, **kw),
                nmatches=mk((max(F - 1, 1),), dtype=torch.int32, **kw))

        self.d_out = [outs(False) for _ in range(self.depth)]
        self.h_out = [outs(True) for _ in range(self.depth)]
        self.done = [None] * self.depth   # last D2H of the step that used the slot
        self.n_submitted = 0
        self.h2d_bytes_per_step = self.n_cams * F * self.H * self.W
        self.d2h_bytes_per_step = sum(t.numel() * t.element_size()
                                      for t in self.h_out[0].kps + self.h_out[0].desc + self.h_out[0].counts) + \
            (self.h_out[0].matches12.numel() + self.h_out[0].nmatches.numel()) * 4 * int(self.match and F > 1)

    @property
    def launch_count(self) -> int:
        return sum(e.launch_count for e in self.ex) + self.matcher.launch_count

    def submit(self, h_images: Sequence) -> int:
        """h_images: one pinned uint8 tensor [F, H, W] per camera.  Returns the step's ticket."""
        torch = self.torch
        step = self.n_submitted
        self.n_submitted += 1
        slot = step % self.depth
        d, h = self.d_out[slot], self.h_out[slot]
        if self.done[slot] is not None:
            # the slot's previous results must have left the device before they are overwritten
            self.s_compute.wait_event(self.done[slot])
        ev_match = None
        for ci in range(self.n_chunks):
            f0, f1 = ci * self.chunk, min(self.F, (ci + 1) * self.chunk)
            if f0 >= f1:
                break
            n = f1 - f0
            b = (step * self.n_chunks + ci) % self.n_ring
            # per-camera events: camera 0 starts computing while camera 1 is still uploading, and its
            # features leave the device while camera 1 computes
            ev_in, ev_done = [], []
            with torch.cuda.stream(self.s_in):
                if self.img_free[b] is not None:
                    self.s_in.wait_event(self.img_free[b])
                for c in range(self.n_cams):
                    self.img[b][c][:n].copy_(h_images[c][f0:f1], non_blocking=True)
                    ev_in.append(torch.cuda.Event())
                    ev_in[c].record(self.s_in)
            with torch.cuda.stream(self.s_compute):
                for c in range(self.n_cams):
                    self.s_compute.wait_event(ev_in[c])
                    self.ex[c].extract_batch_device(self.img[b][c][:n], d.kps[c][f0:f1], d.desc[c][f0:f1], d.counts[c][f0:f1])
                    ev_done.append(torch.cuda.Event())
                    ev_done[c].record(self.s_compute)
                    if c == 0 and f1 == self.F and self.match and self.F > 1:
                        ev_match = self._launch_match(d)
                self.img_free[b] = ev_done[-1]
            with torch.cuda.stream(self.s_out):
                for c in range(self.n_cams):
                    self.s_out.wait_event(ev_done[c])
                    h.kps[c][f0:f1].copy_(d.kps[c][f0:f1], non_blocking=True)
                    h.desc[c][f0:f1].copy_(d.desc[c][f0:f1], non_blocking=True)
                    h.counts[c][f0:f1].copy_(d.counts[c][f0:f1], non_blocking=True)
        if self.match and self.F > 1:
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(ev_match)
                h.matches12.copy_(d.matches12, non_blocking=True)
                h.nmatches.copy_(d.nmatches, non_blocking=True)
        self.done[slot] = torch.cuda.Event()
        self.done[slot].record(self.s_out)
        return step

    def _launch_match(self, d: RigStepResult):
        """Camera 0 of the whole step is extracted (s_compute): match it on the side stream."""
        torch = self.torch
        ev0 = torch.cuda.Event()
        ev0.record(self.s_compute)
        with torch.cuda.stream(self.s_match):
            self.s_match.wait_event(ev0)
            # pairs (t, t+1) of camera 0: the F2 arrays are the same buffers shifted by one frame
            self.matcher.search_for_initialization_device(
                self.F - 1, self.caps[0], d.kps[0], d.desc[0], d.counts[0], d.kps[0][1:], d.desc[0][1:], d.counts[0][1:],
                self.bounds, None, self.window, d.matches12, d.nmatches)
            ev = torch.cuda.Event()
            ev.record(self.s_match)
        return ev

    def result(self, ticket: int) -> RigStepResult:
        """Blocks until the step's results are in pinned host memory."""
        if ticket < self.n_submitted - self.depth or ticket >= self.n_submitted:
            raise ValueError("ticket no longer (or not yet) held by the pipeline")
        slot = ticket % self.depth
        self.done[slot].synchronize()
        return self.h_out[slot]

    def run(self, h_images: Sequence) -> RigStepResult:
        """One synchronous step (submit + result)."""
        return self.result(self.submit(h_images))

    def drain(self) -> None:
        for s in (self.s_in, self.s_compute, self.s_match, self.s_out):
            s.synchronize()
