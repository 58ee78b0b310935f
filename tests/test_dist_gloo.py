"""world_size-2 gloo test (CPU) of the multi-GPU host logic: sharding arithmetic and the
all-gather of per-camera descriptor blocks."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_orb_slam_b200.dist import (allgather_camera_blocks, camera_owner, cameras_of, cross_camera_pairs,
                                       shard_range)


def test_shard_range_partitions():
    for n in (0, 1, 7, 256, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_camera_dealing():
    assert [camera_owner(c, 8) for c in range(8)] == list(range(8))
    assert cameras_of(1, 8, 2) == [1, 3, 5, 7]
    assert cross_camera_pairs(3) == [(0, 1), (1, 2), (2, 0)]


def test_rotated_camera_deal():
    """The deal rotates by one camera per chunk: at every rotation the ranks still partition the cameras, every rank's
    blocks stay contiguous (its all-gather send slot), and over `world` chunks every rank has held every camera."""
    from multi_orb_slam_b200.dist import RigLayout
    n_cams, world = 8, 4
    L = RigLayout(n_cams, world, 5, 7)
    seen = {r: set() for r in range(world)}
    for rot in range(world):
        owned = [cameras_of(r, n_cams, world, rot) for r in range(world)]
        assert sorted(c for o in owned for c in o) == list(range(n_cams))
        for r in range(world):
            assert all(camera_owner(c, world, rot) == r for c in owned[r])
            assert sorted(L.block_of(c, rot) for c in owned[r]) == [r * L.per + j for j in range(L.per)]
            seen[r].update(owned[r])
        assert sorted(L.block_of(c, rot) for c in range(n_cams)) == list(range(n_cams))
        buf = torch.zeros(L.total_bytes, dtype=torch.uint8)
        for c in range(n_cams):
            L.views(buf, c, rot)[2][:] = c + 1
        t = L.match_tables([(1, 6)], 0, 2, rot)
        assert buf.numpy()[t[0, 0]] == 2 and buf.numpy()[t[1, 1]] == 7
    assert all(seen[r] == set(range(n_cams)) for r in range(world))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_cams, F, cap, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cams = cameras_of(rank, n_cams, world)
        g = torch.Generator().manual_seed(0)
        full_desc = torch.randint(0, 256, (n_cams, F, cap, 32), dtype=torch.uint8, generator=g)
        full_kps = torch.rand((n_cams, F, cap, 6), generator=g)
        full_counts = torch.randint(0, cap + 1, (n_cams, F), dtype=torch.int32, generator=g)
        counts, kps, desc = allgather_camera_blocks(full_counts[cams], full_kps[cams], full_desc[cams], n_cams)
        ok = torch.equal(counts, full_counts) and torch.equal(kps, full_kps) and torch.equal(desc, full_desc)
        lo, hi = shard_range(F, rank, world)
        ret[rank] = (bool(ok), lo, hi)
    finally:
        dist.destroy_process_group()


def test_allgather_camera_blocks_world2():
    world, n_cams, F, cap = 2, 4, 3, 5
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n_cams, F, cap, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert ret[0][0] and ret[1][0]
        assert (ret[0][1], ret[0][2], ret[1][1], ret[1][2]) == (0, 2, 2, 3)


def test_rig_layout_blocks_and_tables():
    from multi_orb_slam_b200.dist import RigLayout
    L = RigLayout(8, 4, 5, 7)
    assert L.per == 2 and L.bytes_per_rank == 2 * L.block_bytes and L.total_bytes == 8 * L.block_bytes
    # rank r's cameras r, r+4 occupy blocks 2r, 2r+1: its send slot is contiguous
    assert [L.block_of(c) for c in range(8)] == [0, 2, 4, 6, 1, 3, 5, 7]
    for off in (L.off_kps, L.off_desc, L.block_bytes):
        assert off % 256 == 0
    buf = torch.zeros(L.total_bytes, dtype=torch.uint8)
    for c in range(8):
        counts, kps, desc = L.views(buf, c)
        assert counts.shape == (5,) and kps.shape == (5, 7, 6) and desc.shape == (5, 7, 32)
        counts[:] = 100 * c + torch.arange(5, dtype=torch.int32)
        desc[:] = c
    pairs = cross_camera_pairs(8)
    t = L.match_tables(pairs, 1, 4)
    assert t.shape == (4, 8 * 3)
    raw = buf.numpy()
    for pi, (a, b) in enumerate(pairs):
        for f in range(1, 4):
            p = pi * 3 + (f - 1)
            assert raw[t[0, p]] == a and raw[t[1, p]] == b                      # descriptor rows of the right cameras
            assert raw[t[2, p]:t[2, p] + 4].view(np.int32)[0] == 100 * a + f   # counts of the right frame
            assert raw[t[3, p]:t[3, p] + 4].view(np.int32)[0] == 100 * b + f
            assert t[0, p] % 16 == 0 and t[1, p] % 16 == 0


def _rig_worker(rank, world, port, n_cams, chunk, cap, ret):
    from multi_orb_slam_b200.dist import RigGather, RigLayout
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L = RigLayout(n_cams, world, chunk, cap)
        g = torch.Generator().manual_seed(1)
        full = torch.randint(0, 256, (L.total_bytes,), dtype=torch.uint8, generator=g)  # what every rank must end with
        buf = torch.zeros(L.total_bytes, dtype=torch.uint8)
        for c in cameras_of(rank, n_cams, world):  # "extraction" fills only the rank's own camera blocks
            o = L.block_offset(c)
            buf[o:o + L.block_bytes] = full[o:o + L.block_bytes]
        RigGather(rank, world, backend="torch").allgather_inplace(buf, L.bytes_per_rank)
        ret[rank] = bool(torch.equal(buf, full))
    finally:
        dist.destroy_process_group()


def test_rig_gather_inplace_world2():
    world, n_cams, chunk, cap = 2, 4, 3, 5
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_rig_worker, args=(r, world, port, n_cams, chunk, cap, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
            assert p.exitcode == 0
        assert ret[0] and ret[1]
