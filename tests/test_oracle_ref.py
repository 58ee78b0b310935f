"""Restated extractor (oracle/orb_oracle.cc) vs the reference's ORBextractor.cc compiled verbatim
(oracle/_ref/liborb_ref.so, built only where /root/reference is mounted): identical keypoints,
order, descriptors and pyramids over seeds and geometries.  Also: the verbatim build is
deterministic thanks to the monotonic allocator."""
import numpy as np
import pytest

from multi_orb_slam_b200.synth import textured


@pytest.fixture(scope="module")
def ref_available(oracle_port):
    if oracle_port.load("ref") is None:
        pytest.skip("oracle/_ref/liborb_ref.so not present (no /root/reference on this machine)")
    return oracle_port


@pytest.mark.parametrize("size,nf,seeds", [((640, 480), 1000, range(6)), ((640, 480), 500, range(2)),
                                           ((1241, 376), 2000, range(2)), ((1280, 720), 1000, range(1)),
                                           ((320, 240), 150, range(3))])
def test_port_equals_ref(ref_available, size, nf, seeds):
    O = ref_available
    port, ref = O.extractor("port", nfeatures=nf), O.extractor("ref", nfeatures=nf)
    for seed in seeds:
        img = textured(size[0], size[1], 40 + seed)
        k1, d1, c1 = port.extract(img)
        k2, d2, c2 = ref.extract(img)
        assert np.array_equal(c1, c2)
        assert k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2), (size, nf, seed)
        for l in range(8):
            assert np.array_equal(port.pyramid_level(l), ref.pyramid_level(l))


def test_ref_is_deterministic(ref_available):
    img = textured(640, 480, 77)
    ref = ref_available.extractor("ref")
    k1, d1, _ = ref.extract(img)
    junk = [np.zeros(n) for n in (10, 1000, 100000)]  # perturb the process heap between calls
    k2, d2, _ = ref.extract(img)
    del junk
    assert k1.tobytes() == k2.tobytes() and np.array_equal(d1, d2)


def test_scale_tables_match(ref_available):
    O = ref_available
    for a, b in zip(O.extractor("port").scale_tables(), O.extractor("ref").scale_tables()):
        assert np.array_equal(a, b)
