"""Multi-GPU path (needs >= 2 GPUs): camera streams dealt over ranks, extraction per rank, NCCL
all-gather of the descriptor blocks, cross-camera brute-force matching of each rank's rig-frames;
every rank's result is checked against the CPU oracle."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_cams, F, ret):
    import torch
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import oracle_lib as O
    from multi_orb_slam_b200.dist import allgather_camera_blocks, cameras_of, cross_camera_pairs, shard_range
    from multi_orb_slam_b200.extractor import ORBextractor
    from multi_orb_slam_b200.matcher import ORBmatcher
    from multi_orb_slam_b200.synth import camera_sequence
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        cams = cameras_of(rank, n_cams, world)
        seqs = {c: camera_sequence(320, 240, F, 50 + c) for c in range(n_cams)}  # every rank can rebuild any stream
        ex = ORBextractor(300, 1.2, 8, 20, 7, image_size=(320, 240), max_batch=F, device=rank)
        outs = [ex.extract_batch_device(torch.from_numpy(seqs[c]).cuda()) for c in cams]
        ex.sync()
        kps = torch.stack([o[0].clone() for o in outs]) if len(outs) > 1 else outs[0][0][None].clone()
        counts = torch.stack([o[2] for o in outs])
        # extract_batch_device reuses nothing between calls (outputs allocated per call), so stacking is safe
        desc = torch.stack([o[1] for o in outs])
        kps = torch.stack([o[0] for o in outs])
        counts, kps, desc = allgather_camera_blocks(counts, kps, desc, n_cams)
        lo, hi = shard_range(F, rank, world)
        m = ORBmatcher(0.9, True, device=rank)
        pairs = cross_camera_pairs(n_cams)
        a = torch.tensor([p[0] for p in pairs], device="cuda")
        b = torch.tensor([p[1] for p in pairs], device="cuda")
        cap = desc.shape[2]
        q = desc[a][:, lo:hi].reshape(-1, cap, 32).contiguous()
        t = desc[b][:, lo:hi].reshape(-1, cap, 32).contiguous()
        nq = counts[a][:, lo:hi].reshape(-1).contiguous()
        nt = counts[b][:, lo:hi].reshape(-1).contiguous()
        P = q.shape[0]
        idx, d1, d2 = (torch.empty((P, cap), dtype=torch.int32, device="cuda") for _ in range(3))
        m.bruteforce_batch_device(q, nq, t, nt, idx, d1, d2)
        m.sync()
        idx = idx.cpu().numpy()
        ok = True
        port_ex = O.extractor("port", nfeatures=300)
        for pi, (ca, cb) in enumerate(pairs):
            for f in range(lo, hi):
                ka, da, _ = port_ex.extract(seqs[ca][f])
                kb, db, _ = port_ex.extract(seqs[cb][f])
                ridx, _, _ = O.bruteforce(da, db, 0.9, 50)
                row = pi * (hi - lo) + (f - lo)
                ok = ok and np.array_equal(idx[row, : len(da)], ridx)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_rig_allgather_cross_camera_matching():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world, n_cams, F = 2, 4, 4
    ctx = mp.get_context("spawn")
    port = _free_port()
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, n_cams, F, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(600)
            assert p.exitcode == 0
        assert all(ret[r] for r in range(world))
