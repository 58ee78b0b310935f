"""Dev tool: compare library build variants (tools/dev/liborb_*.so) on extraction stage times."""
import sys, os, glob, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import multi_orb_slam_b200._lib as L
import multi_orb_slam_b200.extractor as E
from multi_orb_slam_b200.synth import camera_sequence
F = 128
imgs = camera_sequence(640, 480, F, 0)
ref = None
for path in sorted(sys.argv[1:] or glob.glob(os.path.join(ROOT, "tools/dev/liborb_nt*.so"))):
    lib = C.CDLL(path)
    for name, (res, args) in L._SIGS.items():
        fn = getattr(lib, name); fn.restype, fn.argtypes = res, args
    L.lib = lib; E.lib = lib
    ex = E.ORBextractor(1000, 1.2, 8, 20, 7, image_size=(640, 480), max_batch=F)
    for _ in range(2):
        out = ex.extract_batch(imgs)
    ex.set_profiling(True)
    for _ in range(5):
        out = ex.extract_batch(imgs)
    ms, n = ex.stage_times_ms()
    sig = (out[2].sum(), out[1][:, :500].sum())
    if ref is None: ref = sig
    print(os.path.basename(path), dict(zip(ex.STAGES, np.round(ms / n, 4))), "same" if sig == ref else "DIFFERENT OUTPUT")
    ex.close()
