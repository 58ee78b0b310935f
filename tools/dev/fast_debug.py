"""Dev probe: FAST candidate lists of the GPU against the oracle, with a classification of the differences."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from multi_orb_slam_b200.extractor import ORBextractor
from multi_orb_slam_b200.synth import textured

O.build_oracle()
size = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 480)
img = textured(size[0], size[1], 0)
port = O.extractor("port", image_size=size) if False else O.extractor("port")
k_ref, d_ref, c_ref = port.extract(img)
ex = ORBextractor(1000, 1.2, 8, 20, 7, image_size=size)
k, d = ex(img)
for l in range(8):
    gx, gy, gs = ex.debug_candidates(l)
    rx, ry, rs = port.candidates(l)
    g = list(zip(gx.tolist(), gy.tolist(), gs.tolist())); r = list(zip(rx.tolist(), ry.tolist(), rs.tolist()))
    sg, sr = set(g), set(r)
    print(f"level {l}: gpu {len(g)} ref {len(r)} only-gpu {len(sg - sr)} only-ref {len(sr - sg)} dup-gpu {len(g) - len(sg)} order-equal {g == r}")
    og = sorted(sg - sr, key=lambda t: (t[1], t[0]))
    rpos = {(x, y): s for x, y, s in r}
    gpos = {(x, y): s for x, y, s in g}
    n_adj = 0
    for x, y, s in og:
        nb = [(dx, dy, rpos[(x + dx, y + dy)]) for dx in (-1, 0, 1) for dy in (-1, 0, 1) if (x + dx, y + dy) in rpos]
        if nb: n_adj += 1
    print(f"   only-gpu with a ref candidate in the 3x3 neighbourhood: {n_adj}")
    print("   only-gpu sample (x,y,score):", og[:12])
    print("   only-gpu y histogram (y mod 7):", np.bincount(np.array([t[1] for t in og], dtype=int) % 7, minlength=7).tolist() if og else [])
    print("   only-ref sample:", sorted(sr - sg, key=lambda t: (t[1], t[0]))[:12])
    if l >= 1: break
