"""GPU parity test of the streaming front end (the C-ABI pipeline orbp_* behind multi_orb_slam_b200/pipeline.py): every step's
keypoints, descriptors and SearchForInitialization matches, delivered to pinned host memory through
the four-stream pipeline, against the CPU oracle; several steps in flight reuse the buffer slots."""
import numpy as np
import pytest

from multi_orb_slam_b200.synth import camera_sequence

pytestmark = pytest.mark.gpu


def _kp(a):
    from multi_orb_slam_b200._lib import KP_DTYPE
    return np.ascontiguousarray(a).view(KP_DTYPE).reshape(-1)


@pytest.mark.parametrize("n_chunks,depth", [(1, 2), (1, 3)])
def test_pipeline_steps_match_oracle(oracle_port, n_chunks, depth):
    import torch
    from multi_orb_slam_b200.pipeline import RigPipeline
    O = oracle_port
    F, W, H, steps = 5, 320, 240, 5
    pipe = RigPipeline((300, 150), 1.2, 8, 20, 7, image_size=(W, H), rig_frames=F, n_chunks=n_chunks, depth=depth,
                       window=100, nnratio=0.9, device=0)
    seqs = [[torch.from_numpy(camera_sequence(W, H, F, 10 * s + c)).pin_memory() for c in range(2)] for s in range(steps)]
    ports = [O.extractor("port", nfeatures=300), O.extractor("port", nfeatures=150)]
    got = {}
    def take(t):
        r = pipe.result(t)
        got[t] = ([x.numpy().copy() for x in r.kps], [x.numpy().copy() for x in r.desc],
                  [x.numpy().copy() for x in r.counts], r.matches12.numpy().copy(), r.nmatches.numpy().copy())

    lag = depth - 1
    for s in range(steps):
        t = pipe.submit(seqs[s])
        if s >= lag:  # consume step s-lag while the later steps are in flight
            take(t - lag)
    for t in range(steps - lag, steps):
        take(t)
    with pytest.raises(ValueError):
        pipe.result(0)  # slot already reused
    for s in range(steps):
        kps, desc, counts, m12, nm = got[s]
        ref = [[ports[c].extract(seqs[s][c][f].numpy())[:2] for f in range(F)] for c in range(2)]
        for c in range(2):
            for f in range(F):
                k_ref, d_ref = ref[c][f]
                n = counts[c][f]
                assert n == len(k_ref), f"step {s} cam {c} frame {f}: count"
                k = _kp(kps[c][f, :n])
                for fld in ("x", "y", "size", "response", "octave"):
                    assert np.array_equal(k[fld], k_ref[fld]), f"step {s} cam {c} frame {f}: {fld}"
                np.testing.assert_allclose(k["angle"], k_ref["angle"], rtol=1e-4, atol=0)
                assert np.array_equal(desc[c][f, :n], d_ref)
        for f in range(F - 1):
            (k1, d1), (k2, d2) = ref[0][f], ref[0][f + 1]
            prev = np.stack([k1["x"], k1["y"]], axis=1).astype(np.float32)
            rn, rm12, _ = O.search_for_initialization(k1, d1, k2, d2, (0, W, 0, H), prev, 100, 0.9, True)
            assert nm[f] == rn, f"step {s} pair {f}: nmatches"
            assert np.array_equal(m12[f, : len(k1)], rm12), f"step {s} pair {f}: matches12"
    assert pipe.launch_count > 0
