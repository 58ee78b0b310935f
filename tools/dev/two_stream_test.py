"""Dev experiment: both cameras' extractors on one stream vs on two streams (fork/join by events)."""
import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from multi_orb_slam_b200.extractor import ORBextractor
from multi_orb_slam_b200.synth import camera_sequence
F, W, H = 256, 640, 480
dev = torch.device("cuda", 0)
imgs = [torch.from_numpy(camera_sequence(W, H, F, s)).to(dev) for s in (0, 1)]
ex = [ORBextractor(nf, 1.2, 8, 20, 7, image_size=(W, H), max_batch=F, device=0) for nf in (1000, 500)]
outs = [(torch.empty((F, e.capacity, 6), device=dev), torch.empty((F, e.capacity, 32), dtype=torch.uint8, device=dev),
         torch.empty((F,), dtype=torch.int32, device=dev)) for e in ex]
s0, s1 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

def run(two):
    ex[0].set_stream(s0.cuda_stream)
    ex[1].set_stream((s1 if two else s0).cuda_stream)
    def step():
        if two:
            ev = torch.cuda.Event(); ev.record(s0); s1.wait_event(ev)
        ex[0].extract_batch_device(imgs[0], *outs[0])
        ex[1].extract_batch_device(imgs[1], *outs[1])
        if two:
            ev2 = torch.cuda.Event(); ev2.record(s1); s0.wait_event(ev2)
    for _ in range(3): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s0)
    for _ in range(10): step()
    b.record(s0)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 10

for two in (False, True, False, True):
    print("two streams" if two else "one stream ", round(run(two), 4), "ms per step (extraction of 512 camera-frames)")
