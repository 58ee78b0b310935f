"""Dev probe: per-stage device time of the extractor on the configs[1] workload (256 frames of 640x480 per launch
group, both nFeatures settings), after a load-based warm-up.  Used for kernel A/B runs:
    ORB_B200_LIB=build/variants/libX.so python tools/dev/stage_times.py [tag]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from multi_orb_slam_b200.extractor import ORBextractor
from multi_orb_slam_b200.synth import camera_sequence

tag = sys.argv[1] if len(sys.argv) > 1 else os.path.basename(os.environ.get("ORB_B200_LIB", "default"))
W, H, F = (int(x) for x in (sys.argv[2:5] if len(sys.argv) > 4 else (640, 480, 256)))
st = torch.cuda.Stream()
imgs = [torch.from_numpy(camera_sequence(W, H, min(F, 32), c)).cuda() for c in range(2)]
imgs = [torch.cat([i] * (F // i.shape[0]))[:F].contiguous() for i in imgs]
exs = [ORBextractor(nf, 1.2, 8, 20, 7, image_size=(W, H), max_batch=F, device=0) for nf in (1000, 500)]
outs = []
for e, im in zip(exs, imgs):
    e.set_stream(st.cuda_stream)
    outs.append(e.extract_batch_device(im))
t0 = time.perf_counter()
while time.perf_counter() - t0 < 0.5:
    for e, im, o in zip(exs, imgs, outs):
        e.extract_batch_device(im, *o)
    st.synchronize()
for e in exs:
    e.set_profiling(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
K = 10
e0.record(st)
for _ in range(K):
    for e, im, o in zip(exs, imgs, outs):
        e.extract_batch_device(im, *o)
e1.record(st)
st.synchronize()
tot = np.zeros(5)
for e in exs:
    s, n = e.stage_times_ms()
    tot += s / K
chk = int(sum(int(o[2].sum().item()) for o in outs))
print(f"{tag}: step {e0.elapsed_time(e1) / K:.3f} ms | pyramid {tot[0]:.3f} fast {tot[1]:.3f} octree {tot[2]:.3f} blur {tot[3]:.3f} orient {tot[4]:.3f} | keypoints {chk}")
