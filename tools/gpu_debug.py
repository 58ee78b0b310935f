"""First-contact debugging on the GPU box: per-stage mismatch report against the oracle."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from multi_orb_slam_b200.extractor import ORBextractor
from multi_orb_slam_b200.synth import textured

O.build_oracle()
size = (640, 480)
img = textured(size[0], size[1], 0)
port = O.extractor("port")
k_ref, d_ref, c_ref = port.extract(img)
ex = ORBextractor(1000, 1.2, 8, 20, 7, image_size=size)
k, d = ex(img)
print("counts", len(k), len(k_ref), c_ref)
for l in range(8):
    a, b = ex.pyramid_level(l, with_border=True), port.pyramid_level(l)
    print("pyr", l, a.shape, b.shape, "diff px:", int((a != b).sum()) if a.shape == b.shape else "shape")
for l in range(8):
    gx, gy, gs = ex.debug_candidates(l)
    rx, ry, rs = port.candidates(l)
    same = len(gx) == len(rx) and np.array_equal(gx, rx) and np.array_equal(gy, ry) and np.array_equal(gs, rs)
    print("cand", l, len(gx), len(rx), "same" if same else "DIFF")
    if not same:
        sg = set(zip(gx.tolist(), gy.tolist(), gs.tolist())); sr = set(zip(rx.tolist(), ry.tolist(), rs.tolist()))
        print("   only gpu:", sorted(sg - sr)[:8], " only ref:", sorted(sr - sg)[:8], "set-equal:", sg == sr)
for l in range(8):
    lv = ex.pyramid_level(l)
    rb = port.blurred(l, lv.shape[1], lv.shape[0])
    gb = ex.debug_blurred(l)
    print("blur", l, "diff px:", int((gb != rb).sum()) if rb is not None else "n/a")
n = min(len(k), len(k_ref))
for f in ("x", "y", "octave", "response", "size", "angle"):
    print("kp", f, "mismatch:", int((k[f][:n] != k_ref[f][:n]).sum()))
print("desc rows differ:", int((d[:n] != d_ref[:n]).any(axis=1).sum()))
