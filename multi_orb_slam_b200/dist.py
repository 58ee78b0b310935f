"""Multi-GPU plumbing: one process per GPU (torch.distributed; NCCL over NVLink on the GPUs, gloo
in the CPU tests).  The reference is single-process and single-device; what
shards here are its independent units — camera-frames for extraction, rig-frames for matching.
The ONLY collective is the all-gather of per-camera-frame keypoint/descriptor blocks that gives
every rank the descriptors of all cameras (the multi-GPU analogue of Frame::mDescriptors_total,
src/Frame.cc:170,191-194) before cross-camera matching."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous balanced block of [0, n_units) owned by `rank` (first n_units % world ranks get one more)."""
    base, extra = divmod(n_units, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def camera_owner(cam: int, world: int) -> int:
    """Config 5: camera streams are dealt round-robin, so 8 GPUs take one camera each."""
    return cam % world


def cameras_of(rank: int, n_cams: int, world: int) -> List[int]:
    return [c for c in range(n_cams) if camera_owner(c, world) == rank]


def cross_camera_pairs(n_cams: int) -> List[Tuple[int, int]]:
    """Camera c is matched against camera (c+1) mod n_cams (ring of overlapping fields of view)."""
    return [(c, (c + 1) % n_cams) for c in range(n_cams)]


def allgather_camera_blocks(counts: torch.Tensor, kps: torch.Tensor, desc: torch.Tensor, n_cams: int, group=None):
    """All-gather fixed-stride per-camera-frame blocks.

    Each rank holds, for its cameras (cameras_of(rank)), [n_local_cams, F, ...] blocks: counts
    [n_local_cams, F] i32, kps [n_local_cams, F, cap, 6] f32, desc [n_local_cams, F, cap, 32] u8.
    Returns the same three tensors for ALL cameras, indexed by camera id: [n_cams, F, ...].
    Requires n_cams % world == 0 (equal block sizes, one all_gather_into_tensor per tensor)."""
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
