"""Dev tool: dump the oracle's FAST candidates of one level as octree_bench input."""
import sys, os, struct
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, oracle_lib as O
from multi_orb_slam_b200.synth import textured
level = int(sys.argv[2]) if len(sys.argv) > 2 else 0
port = O.extractor("port"); port.extract(textured(640, 480, 0))
x, y, s = port.candidates(level)
pyr = port.pyramid_level(level); lh, lw = pyr.shape[0] - 38, pyr.shape[1] - 38
N = int(port.features_per_level()[level])
keys = (x.astype(np.uint32) | (y.astype(np.uint32) << 12) | (s.astype(np.uint32) << 24)).astype("<u4")
with open(sys.argv[1], "wb") as f:
    f.write(struct.pack("<4i", len(keys), lw - 32, lh - 32, N)); f.write(keys.tobytes())
print(len(keys), lw - 32, lh - 32, N)
