#!/usr/bin/env python
"""Golden vectors of the reference's OWN matcher (src/ORBmatcher.cc compiled verbatim, oracle/_ref/libmatcher_ref.so)
for tests/test_golden.py: outputs only; the inputs are the seeded scenes of tests/test_gpu_matcher.py, rebuilt by the
test.  Run where /root/reference exists (make -C oracle first).  Writes tests/golden/ref_matcher.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as O  # noqa: E402
from golden_matcher_cases import CASES  # noqa: E402

assert O.load("mref") is not None, "oracle/_ref/libmatcher_ref.so missing: run `make -C oracle` where /root/reference exists"
out = {}
for name, run in CASES.items():
    res = run(O, "ref")
    for j, a in enumerate(res):
        out[f"{name}_{j}"] = np.asarray(a)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_matcher.npz"), **out)
print("wrote ref_matcher.npz:", ", ".join(sorted(out)))
