"""Front end of a many-camera rig sharded over the GPUs of one box (BASELINE.json configs[4]: 8 cameras).

Reference analogue: Frame::Frame runs one ORBextractor per camera and concatenates the cameras' descriptors into
mDescriptors_total (src/Frame.cc:148-346, :170, :191-194); the matchers then loop over the cameras of that matrix
(src/ORBmatcher.cc:628, 2030, 2269, 3582).  Here, per rank (one process per GPU):

    extract   the rank's cameras (dealt round-robin, dist.camera_owner; the deal rotates by one camera per chunk so
              that every rank sees every camera in turn: the cameras' textures, hence costs, differ by a few percent
              and a fixed deal would make the busiest camera's rank the pace of the step), a chunk of rig-frames at
              a time; the extractor's last kernel writes counts / keypoints / descriptors STRAIGHT into the rank's
              slot of the chunk's gather buffer (dist.RigLayout) — no staging copy, one buffer, one collective
    gather    ONE in-place all-gather per chunk on its own stream (NCCL through the C ABI, orbd_allgather_inplace):
              it runs underneath the extraction of the next chunk
    match     cross-camera brute-force matching (ring of camera pairs) of the rank's share of the chunk's rig-frames,
              all pairs of the chunk in one launch (orbm_bruteforce_indexed_device), on a third stream

A step covers `rig_frames` rig-frames of all cameras.  world == 1 runs the same code without the collective.
There is no CPU fallback: extraction and matching are the CUDA library's."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .dist import RigGather, RigLayout, cameras_of, cross_camera_pairs, shard_range
from .extractor import ORBextractor
from .matcher import ORBmatcher


@dataclass
class RigStepResult:
    """Device tensors of one step, valid until the next step() of the same RigFrontEnd.

    idx / d1 / d2: [n_chunks][n_cam_pairs * shard_len, cap] i32 — best target of every query keypoint of camera a
    in camera b (-1 = rejected by TH_LOW / ratio), best and second-best distance; row = pi*shard_len + (f - lo) for
    camera pair pi and rig-frame f of the chunk's shard [lo, hi).  shards: [(chunk_first_frame, lo, hi)]."""
    idx: list
    d1: list
    d2: list
    shards: List[Tuple[int, int, int]]
    rots: List[int] = None            # rotation of the camera deal in every chunk (RigLayout.views(buf, cam, rot))
    collected: Optional[list] = None  # with collect=True: a copy of every chunk's gather buffer (tests)


class RigFrontEnd:
    def __init__(self, n_cams: int = 8, nfeatures: int = 1000, scaleFactor: float = 1.2, nlevels: int = 8,
                 iniThFAST: int = 20, minThFAST: int = 7, *, image_size: Tuple[int, int] = (1280, 720),
                 rig_frames: int = 512, chunk: int = 64, rank: int = 0, world: int = 1, device: int = 0,
                 nnratio: float = 0.9, th_dist: int = 50, depth: int = 2, backend: str = "orbd", group=None,
                 pairs: Optional[Sequence[Tuple[int, int]]] = None, rotate: bool = True):
        import torch
        self.torch = torch
        self.n_cams, self.F, self.rank, self.world = int(n_cams), int(rig_frames), int(rank), int(world)
        self.chunk = max(1, min(int(chunk), self.F))
        self.n_chunks = (self.F + self.chunk - 1) // self.chunk
        self.depth = max(2, int(depth))
        self.W, self.H = int(image_size[0]), int(image_size[1])
        self.dev = torch.device("cuda", device)
        self.rotate = bool(rotate) and self.world > 1
        self.cams = cameras_of(self.rank, self.n_cams, self.world)  # the rank's cameras in chunk 0 (all chunks if not rotating)
        # every camera whose frames this rank reads during a step (images[c] must exist for these)
        self.input_cams = sorted({c for k in range(self.n_chunks) for c in self.cams_of_chunk(k)})
        self.pairs = list(pairs) if pairs is not None else cross_camera_pairs(self.n_cams)
        self.nnratio, self.th_dist = float(nnratio), int(th_dist)
        self.ex = ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, image_size=image_size,
                               max_batch=self.chunk, device=device)
        self.cap = self.ex.capacity
        self.layout = RigLayout(self.n_cams, self.world, self.chunk, self.cap)
        self.gather = RigGather(self.rank, self.world, backend=backend, device=device, group=group)
        self.matcher = ORBmatcher(nnratio, True, device=device)
        # the collective and the matcher are small next to the extraction kernels that fill the GPU: high priority, so
        # their CTAs are placed as soon as SMs drain instead of queueing behind whole extraction grids
        self.s_compute, self.s_comm = torch.cuda.Stream(device=self.dev), torch.cuda.Stream(device=self.dev, priority=-1)
        self.s_match = torch.cuda.Stream(device=self.dev, priority=-1)
        self.ex.set_stream(self.s_compute.cuda_stream)
        self.matcher.set_stream(self.s_match.cuda_stream)
        # `depth` gather buffers in flight: chunk k+1 is extracted while chunk k travels and chunk k-1 is matched
        self.bufs = [torch.zeros(self.layout.total_bytes, dtype=torch.uint8, device=self.dev) for _ in range(self.depth)]
        self.buf_free = [None] * self.depth  # event: the matcher has finished reading the buffer
        self._tables: Dict[int, tuple] = {}
        self._out: Dict[Tuple[int, int], tuple] = {}
        self.skip_gather = False  # measurement only (bench.py): time the step without its collective
        self.allgather_bytes_per_chunk = self.layout.bytes_per_rank * (self.world - 1) if self.world > 1 else 0

    def rot_of_chunk(self, k: int) -> int:
        return k % self.world if self.rotate else 0

    def cams_of_chunk(self, k: int) -> List[int]:
        return cameras_of(self.rank, self.n_cams, self.world, self.rot_of_chunk(k))

    @property
    def launch_count(self) -> int:
        return self.ex.launch_count + self.matcher.launch_count

    def _match_tables(self, n: int, rot: int):
        """Offset tables (device) of the rank's shard of a chunk holding n rig-frames, camera deal rotated by rot."""
        if (n, rot) not in self._tables:
            lo, hi = shard_range(n, self.rank, self.world)
            t = self.layout.match_tables(self.pairs, lo, hi, rot)
            self._tables[(n, rot)] = (lo, hi, self.torch.from_numpy(t).to(self.dev))
        return self._tables[(n, rot)]

    def step(self, images: Dict[int, "object"], collect: bool = False) -> RigStepResult:
        """images[c]: uint8 CUDA tensor [rig_frames, H, W] for every camera c in self.input_cams (of a camera's
        frames the rank reads only the chunks in which the rotating deal hands it that camera).
        Asynchronous: returns once everything is enqueued; call sync() before reading the result."""
        torch = self.torch
        res = RigStepResult([], [], [], [], [], [] if collect else None)
        cur = torch.cuda.current_stream(self.dev)
        ev0 = torch.cuda.Event()
        ev0.record(cur)
        self.s_compute.wait_event(ev0)  # the caller produced `images` on its current stream
        for k in range(self.n_chunks):
            f0, f1 = k * self.chunk, min(self.F, (k + 1) * self.chunk)
            n = f1 - f0
            rot = self.rot_of_chunk(k)
            b = k % self.depth
            buf = self.bufs[b]
            if self.buf_free[b] is not None:
                self.s_compute.wait_event(self.buf_free[b])
            with torch.cuda.stream(self.s_compute):
                for c in self.cams_of_chunk(k):
                    counts, kps, desc = self.layout.views(buf, c, rot)
                    self.ex.extract_batch_device(images[c][f0:f1], kps[:n], desc[:n], counts[:n])
                ev_x = torch.cuda.Event()
                ev_x.record(self.s_compute)
            if self.world > 1 and not self.skip_gather:
                self.s_comm.wait_event(ev_x)
                with torch.cuda.stream(self.s_comm):
                    self.gather.allgather_inplace(buf, self.layout.bytes_per_rank, self.s_comm.cuda_stream)
                    ev_g = torch.cuda.Event()
                    ev_g.record(self.s_comm)
            else:
                ev_g = ev_x
            lo, hi, tab = self._match_tables(n, rot)
            key = (k, n)
            if key not in self._out:
                rows = len(self.pairs) * (hi - lo)
                self._out[key] = tuple(torch.empty((rows, self.cap), dtype=torch.int32, device=self.dev) for _ in range(3))
            idx, d1, d2 = self._out[key]
            self.s_match.wait_event(ev_g)
            with torch.cuda.stream(self.s_match):
                if hi > lo:
                    self.matcher.bruteforce_indexed_device(buf, tab, idx, d1, d2, self.cap, th_dist=self.th_dist,
                                                           ratio=self.nnratio)
                if collect:
                    res.collected.append(buf.clone())
                ev_m = torch.cuda.Event()
                ev_m.record(self.s_match)
            self.buf_free[b] = ev_m
            res.idx.append(idx); res.d1.append(d1); res.d2.append(d2)
            res.shards.append((f0, lo, hi))
            res.rots.append(rot)
        # the caller's stream continues after the whole step
        ev_end = torch.cuda.Event()
        ev_end.record(self.s_match)
        cur.wait_event(ev_end)
        ev_c = torch.cuda.Event()
        ev_c.record(self.s_compute)
        cur.wait_event(ev_c)
        return res

    def sync(self) -> None:
        for s in (self.s_compute, self.s_comm, self.s_match):
            s.synchronize()

    def close(self) -> None:
        self.sync()
        self.gather.close()
