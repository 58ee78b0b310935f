"""configs[4] code path (multi_orb_slam_b200/rig.py): 8-camera 1280x720 rig, extraction written straight into the
gather buffer, one in-place all-gather per chunk, cross-camera matching of the rank's rig-frame shard — every output
checked against the CPU oracle.  world 1 runs on any GPU box (same code, no collective); world 2 needs two GPUs and
exercises the NCCL all-gather behind the C ABI (orbd_*)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

W, H, NF = 1280, 720, 1000


def _check_rank(rank, world, n_cams, F, chunk, backend="orbd"):
    """Runs one RigFrontEnd step on this process's GPU `rank` and compares with the oracle.  Returns a list of
    mismatch descriptions (empty = parity)."""
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import oracle_lib as O
    from multi_orb_slam_b200._lib import KP_DTYPE
    from multi_orb_slam_b200.rig import RigFrontEnd
    from multi_orb_slam_b200.synth import camera_sequence
    seqs = {c: camera_sequence(W, H, F, 70 + c) for c in range(n_cams)}  # every rank can rebuild any stream
    fe = RigFrontEnd(n_cams, NF, 1.2, 8, 20, 7, image_size=(W, H), rig_frames=F, chunk=chunk, rank=rank, world=world,
                     device=rank, backend=backend)
    images = {c: torch.from_numpy(seqs[c]).to(f"cuda:{rank}") for c in fe.input_cams}
    res = fe.step(images, collect=True)
    fe.sync()
    bad = []
    port = O.extractor("port", nfeatures=NF)
    ref = {}
    for c in range(n_cams):
        for f in range(F):
            ref[(c, f)] = port.extract(seqs[c][f])[:2]
    L = fe.layout
    for k, (f0, lo, hi) in enumerate(res.shards):
        n = min(chunk, F - f0)
        buf = res.collected[k]
        # after the gather every rank holds every camera's block of the chunk
        for c in range(n_cams):
            counts, kps, desc = (t.cpu().numpy() for t in L.views(buf, c, res.rots[k]))
            for j in range(n):
                rk, rd = ref[(c, f0 + j)]
                m = int(counts[j])
                g = np.ascontiguousarray(kps[j, :m]).view(KP_DTYPE).reshape(-1)
                if m != len(rk):
                    bad.append(f"count cam {c} frame {f0 + j}: {m} vs {len(rk)}")
                    continue
                for name in ("x", "y", "octave", "response", "size", "angle"):
                    if not np.array_equal(g[name], rk[name]):
                        bad.append(f"{name} cam {c} frame {f0 + j}")
                if not np.array_equal(desc[j, :m], rd):
                    bad.append(f"descriptors cam {c} frame {f0 + j}")
        idx, d1, d2 = (t.cpu().numpy() for t in (res.idx[k], res.d1[k], res.d2[k]))
        for pi, (a, b) in enumerate(fe.pairs):
            for f in range(lo, hi):
                da, db = ref[(a, f0 + f)][1], ref[(b, f0 + f)][1]
                ri, r1, r2 = O.bruteforce(da, db, 0.9, 50)
                row = pi * (hi - lo) + (f - lo)
                if not (np.array_equal(idx[row, :len(da)], ri) and np.array_equal(d1[row, :len(da)], r1) and
                        np.array_equal(d2[row, :len(da)], r2)):
                    bad.append(f"matches pair ({a},{b}) frame {f0 + f}")
    fe.close()
    return bad, sum(hi - lo for _, lo, hi in res.shards)


def test_rig8_world1_vs_oracle():
    """The whole configs[4] step on ONE GPU: 8 cameras x 3 rig-frames of 1280x720 in chunks of 2 (ragged last chunk,
    gather-buffer reuse), no collective."""
    bad, n_matched = _check_rank(0, 1, 8, 3, 2)
    assert not bad, bad[:10]
    assert n_matched == 3


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    # torch.distributed only carries the 128-byte NCCL id to the other rank; the data path is orbd_allgather_inplace
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bad, n = _check_rank(rank, world, 4, 4, 2)
        ret[rank] = (bad[:10], n)
    finally:
        dist.destroy_process_group()


def test_rig_world2_nccl_allgather_vs_oracle():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    port = _free_port()
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(900)
            assert p.exitcode == 0
        for r in range(world):
            assert not ret[r][0], ret[r][0]
        assert ret[0][1] + ret[1][1] == 4  # the two ranks' shards cover every rig-frame once
