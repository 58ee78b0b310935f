"""Dev tool: run the real extractor built with -DORB_OT_TIMING and print block (0,0)'s octree timeline."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import multi_orb_slam_b200._lib as L
import ctypes as C
import numpy as np
lib = C.CDLL(os.path.join(ROOT, "tools/dev/liborb_b200_timing.so"))
for name, (res, args) in L._SIGS.items():
    fn = getattr(lib, name); fn.restype, fn.argtypes = res, args
L.lib = lib
import multi_orb_slam_b200.extractor as E
E.lib = lib
from multi_orb_slam_b200.synth import textured
F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
imgs = np.stack([textured(640, 480, i % 4) for i in range(F)])
ex = E.ORBextractor(1000, 1.2, 8, 20, 7, image_size=(640, 480), max_batch=F)
for _ in range(3):
    ex.extract_batch(imgs)
ex.close()
