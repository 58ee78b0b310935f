"""Multi-GPU plumbing: one process per GPU.  The reference is single-process and single-device
(SURVEY.md §2.2); what shards here are its independent units — camera streams for extraction,
rig-frames for matching.  The ONLY collective is the all-gather of per-camera keypoint/descriptor
blocks that gives every rank the descriptors of all cameras of a rig-frame (the multi-GPU analogue
of Frame::mDescriptors_total, src/Frame.cc:170,191-194) before cross-camera matching
(src/ORBmatcher.cc:628,3582 loop over the cameras of that matrix).

Data path: the extractor writes counts / keypoints / descriptors of a chunk of rig-frames straight
into this rank's slot of ONE gather buffer (RigLayout), and ONE in-place all-gather per chunk moves it
— through the C ABI (`orbd_allgather_inplace`, NCCL bound inside liborb_b200.so, backend "orbd") on
GPUs, or through torch.distributed (backend "torch": gloo in the CPU tests of the host logic)."""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced block of [0, n_units) owned by `rank` (first n_units % world ranks get one more)."""
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def camera_owner(cam: int, world: int, rot: int = 0) -> int:
    """Config 5: camera streams are dealt round-robin, so 8 GPUs take one camera each.  `rot` rotates the deal
    (RigFrontEnd advances it by one per chunk of rig-frames: the cameras differ in texture, hence in cost, and a rank
    that kept the busiest camera for the whole batch would set the pace for all the others)."""
    return (cam + rot) % world


def cameras_of(rank: int, n_cams: int, world: int, rot: int = 0) -> List[int]:
    return [c for c in range(n_cams) if camera_owner(c, world, rot) == rank]


def cross_camera_pairs(n_cams: int) -> List[Tuple[int, int]]:
    """Camera c is matched against camera (c+1) mod n_cams (ring of overlapping fields of view)."""
    return [(c, (c + 1) % n_cams) for c in range(n_cams)]


def _align(n: int, a: int = 256) -> int:
    return (n + a - 1) // a * a


class RigLayout:
    """Byte layout of the gather buffer of one chunk of `chunk` rig-frames.

    world * per camera blocks, rank-major: block r*per + j belongs to the j-th camera of rank r under the round-robin
    deal of camera_owner (rotated by `rot`), so rank r's `per` blocks are contiguous = its send slot of the in-place
    all-gather.  A block:
    counts [chunk] i32 | keypoints [chunk, cap, 6] f32 (orbx_keypoint rows) | descriptors [chunk, cap, 32] u8,
    sections 256-byte aligned."""

    def __init__(self, n_cams: int, world: int, chunk: int, cap: int):
        if n_cams % world:
            raise ValueError("cameras must divide evenly over the ranks")
        self.n_cams, self.world, self.chunk, self.cap = int(n_cams), int(world), int(chunk), int(cap)
        self.per = self.n_cams // self.world
        self.off_counts = 0
        self.off_kps = _align(4 * self.chunk)
        self.off_desc = self.off_kps + _align(24 * self.chunk * self.cap)
        self.block_bytes = self.off_desc + _align(32 * self.chunk * self.cap)
        self.bytes_per_rank = self.per * self.block_bytes
        self.total_bytes = self.world * self.bytes_per_rank

    def block_of(self, cam: int, rot: int = 0) -> int:
        return camera_owner(cam, self.world, rot) * self.per + cam // self.world

    def block_offset(self, cam: int, rot: int = 0) -> int:
        return self.block_of(cam, rot) * self.block_bytes

    def views(self, buf, cam: int, rot: int = 0):
        """(counts [chunk] i32, kps [chunk, cap, 6] f32, desc [chunk, cap, 32] u8) views of camera `cam`'s block
        inside the flat uint8 torch tensor `buf`."""
        import torch
        o = self.block_offset(cam, rot)
        counts = buf[o + self.off_counts: o + self.off_counts + 4 * self.chunk].view(torch.int32)
        kps = buf[o + self.off_kps: o + self.off_kps + 24 * self.chunk * self.cap].view(torch.float32).view(self.chunk, self.cap, 6)
        desc = buf[o + self.off_desc: o + self.off_desc + 32 * self.chunk * self.cap].view(self.chunk, self.cap, 32)
        return counts, kps, desc

    def match_tables(self, pairs: Sequence[Tuple[int, int]], lo: int, hi: int, rot: int = 0) -> np.ndarray:
        """Offset tables of orbm_bruteforce_indexed_device for rig-frames [lo, hi) of the chunk and the camera pairs
        `pairs`: int64 [4, n] (query rows, target rows, query count, target count), pair index = pi*(hi-lo) + (f-lo)."""
        n = len(pairs) * (hi - lo)
        t = np.zeros((4, n), dtype=np.int64)
        f = np.arange(lo, hi, dtype=np.int64)
        for pi, (a, b) in enumerate(pairs):
            s = slice(pi * (hi - lo), (pi + 1) * (hi - lo))
            oa, ob = self.block_offset(a, rot), self.block_offset(b, rot)
            t[0, s] = oa + self.off_desc + f * (32 * self.cap)
            t[1, s] = ob + self.off_desc + f * (32 * self.cap)
            t[2, s] = oa + self.off_counts + 4 * f
            t[3, s] = ob + self.off_counts + 4 * f
        return t


class RigGather:
    """The in-place all-gather of a RigLayout buffer.

    backend "orbd": NCCL through the C ABI (orbd_*; the communicator is created from a unique id that rank 0
    publishes through torch.distributed's store — host plumbing only); backend "torch": torch.distributed
    collectives (gloo CPU tests, or NCCL for A/B); world 1: no-op."""

    def __init__(self, rank: int, world: int, backend: str = "orbd", device: int = -1, group=None):
        self.rank, self.world, self.backend, self.group = int(rank), int(world), backend, group
        self._comm = None
        if self.world > 1 and backend == "orbd":
            import torch
            import torch.distributed as dist
            from ._lib import OK, OrbError, check_d, lib
            if "ORB_NCCL_LIB" not in os.environ:
                # the process already carries torch's NCCL: bind that one rather than a second copy
                cand = os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    os.environ["ORB_NCCL_LIB"] = cand
            uid = np.zeros(128, dtype=np.uint8)
            if self.rank == 0:
                check_d(None, lib.orbd_get_unique_id(uid.ctypes.data))
            box = [uid.tobytes()]
            dist.broadcast_object_list(box, src=0, group=group)
            uid = np.frombuffer(box[0], dtype=np.uint8).copy()
            h = C.c_void_p()
            rc = lib.orbd_comm_create(self.rank, self.world, uid.ctypes.data, device, C.byref(h))
            if rc != OK:
                raise OrbError(rc, (lib.orbd_last_error(None) or b"").decode())
            self._comm = h

    def close(self) -> None:
        if self._comm is not None:
            from ._lib import lib
            lib.orbd_comm_destroy(self._comm)
            self._comm = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def allgather_inplace(self, buf, bytes_per_rank: int, cuda_stream: int = 0) -> None:
        """buf: flat uint8 torch tensor of world*bytes_per_rank bytes whose slot `rank` is complete (in `cuda_stream`
        order for backend "orbd"; on the current stream for backend "torch").  Asynchronous on GPUs."""
        if self.world == 1:
            return
        if self.backend == "orbd":
            from ._lib import check_d, lib
            check_d(self._comm, lib.orbd_allgather_inplace(self._comm, buf.data_ptr(), bytes_per_rank, cuda_stream or None))
            return
        import torch.distributed as dist
        mine = buf[self.rank * bytes_per_rank:(self.rank + 1) * bytes_per_rank]
        if buf.is_cuda:
            dist.all_gather_into_tensor(buf, mine, group=self.group)
        else:
            parts = [buf[r * bytes_per_rank:(r + 1) * bytes_per_rank] for r in range(self.world)]
            dist.all_gather(parts, mine.clone(), group=self.group)


def allgather_camera_blocks(counts, kps, desc, n_cams: int, group=None):
    """All-gather fixed-stride per-camera-frame blocks held as three separate tensors (round-1 interface, kept for
    callers that do not use a RigLayout buffer).

    Each rank holds, for its cameras (cameras_of(rank)), [n_local_cams, F, ...] blocks: counts
    [n_local_cams, F] i32, kps [n_local_cams, F, cap, 6] f32, desc [n_local_cams, F, cap, 32] u8.
    Returns the same three tensors for ALL cameras, indexed by camera id: [n_cams, F, ...].
    Requires n_cams % world == 0 (equal block sizes, one all_gather_into_tensor per tensor)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return counts, kps, desc
    assert n_cams % world == 0, "cameras must divide evenly over the ranks"
    per = n_cams // world
    assert counts.shape[0] == per

    def gather(x):
        out = torch.empty((world * per,) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        # row r*per + j holds camera r + j*world (round-robin deal): reorder to camera id
        out = out.view((world, per) + tuple(x.shape[1:]))
        return out.transpose(0, 1).reshape((n_cams,) + tuple(x.shape[1:]))

    return gather(counts), gather(kps), gather(desc)
