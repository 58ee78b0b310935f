"""The product's octree policy core (multi_orb_slam_b200/csrc/octree_core.h) compiled for the
host under its sequential thread emulation, checked against the oracle's DistributeOctTree
restatement on real FAST candidates: same survivors, same order.  Validates the parallel
reformulation (rounds, stable slots, early break, tie-break) without a GPU."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from multi_orb_slam_b200.synth import textured

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def host_lib():
    so = os.path.join(HERE, "native", "liboctree_host.so")
    tmp = f"{so}.{os.getpid()}.tmp"  # built aside and renamed: pytest-xdist workers may race here
    subprocess.run(["g++", "-O2", "-fPIC", "-shared", "-o", tmp, os.path.join(HERE, "native", "octree_host.cc")], check=True)
    os.replace(tmp, so)
    return C.CDLL(so)


def _run(lib, x, y, s, W, H, N):
    cap = N + 64
    ox, oy, os_ = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    n = lib.octree_host_distribute(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p), s.ctypes.data_as(C.c_void_p),
                                   len(x), W, H, N, ox.ctypes.data_as(C.c_void_p), oy.ctypes.data_as(C.c_void_p),
                                   os_.ctypes.data_as(C.c_void_p), cap)
    assert n >= 0
    return ox[:n], oy[:n], os_[:n]


@pytest.mark.parametrize("w,h,nf", [(640, 480, 1000), (640, 480, 500), (1241, 376, 2000), (1280, 720, 1000),
                                     (640, 480, 5000), (320, 240, 50), (640, 480, 30000)])
def test_octree_core_matches_oracle(host_lib, oracle_port, w, h, nf):
    port = oracle_port.extractor("port", nfeatures=nf)
    quota = port.features_per_level()
    for seed in range(2):
        port.extract(textured(w, h, seed + 10))
        for l in range(8):
            x, y, s = port.candidates(l)
            pyr = port.pyramid_level(l)
            lh, lw = pyr.shape[0] - 38, pyr.shape[1] - 38
            kx, ky, kr, _ = port.level_keypoints(l)
            ox, oy, os_ = _run(host_lib, x, y, s, lw - 32, lh - 32, int(quota[l]))
            assert len(ox) == len(kx), (w, h, nf, seed, l)
            assert np.array_equal(ox + 16, kx) and np.array_equal(oy + 16, ky) and np.array_equal(os_, kr)


def test_octree_core_degenerate_inputs(host_lib):
    e = np.zeros(0, dtype=np.int32)
    assert len(_run(host_lib, e, e, e, 608, 448, 217)[0]) == 0
    one = np.array([5], dtype=np.int32)
    ox, oy, os_ = _run(host_lib, one, one, np.array([30], dtype=np.int32), 608, 448, 217)
    assert list(ox) == [5] and list(oy) == [5] and list(os_) == [30]
    # two adjacent keys, N=1: one split keeps both nodes (list can overshoot N)
    x = np.array([10, 600], dtype=np.int32); y = np.array([10, 400], dtype=np.int32); s = np.array([9, 8], dtype=np.int32)
    ox, _, _ = _run(host_lib, x, y, s, 608, 448, 1)
    assert sorted(ox) == [10, 600]
