"""Dev probe: extraction time of each of the 8 synthetic rig cameras (seeds 300..307) on one GPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from multi_orb_slam_b200.extractor import ORBextractor
from multi_orb_slam_b200.synth import camera_sequence
F = 128
ex = ORBextractor(1000, 1.2, 8, 20, 7, image_size=(1280, 720), max_batch=F, device=0)
st = torch.cuda.Stream()
ex.set_stream(st.cuda_stream)
for c in range(8):
    base = torch.from_numpy(camera_sequence(1280, 720, 8, 300 + c)).cuda()
    img = torch.empty((F, 720, 1280), dtype=torch.uint8, device="cuda")
    for i in range(0, F, 8):
        img[i:i + 8] = base
    torch.cuda.synchronize()
    out = ex.extract_batch_device(img)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        ex.extract_batch_device(img, *out)
    e1.record(st)
    st.synchronize()
    ex.set_profiling(True)
    ex.extract_batch_device(img, *out)
    s, _ = ex.stage_times_ms()
    ex.set_profiling(False)
    print(f"camera {c}: {e0.elapsed_time(e1) / 5 / F * 1e3:.2f} us/frame, stages {[round(float(x) / F * 1e3, 2) for x in s]}")
