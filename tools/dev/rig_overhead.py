"""Dev probe: per-chunk host overhead of RigFrontEnd.step — the per-rank workload of the 8-GPU run (one camera per
rank) emulated on one GPU (n_cams = 1, self pairs) next to the 8-camera single-GPU case."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from multi_orb_slam_b200.rig import RigFrontEnd
from multi_orb_slam_b200.synth import camera_sequence

F = 512
for n_cams, chunk in ((8, 64), (1, 64), (1, 128), (1, 256)):
    fe = RigFrontEnd(n_cams, 1000, 1.2, 8, 20, 7, image_size=(1280, 720), rig_frames=F, chunk=chunk, rank=0, world=1, device=0,
                     pairs=[(c, (c + 1) % n_cams) for c in range(n_cams)])
    base = torch.from_numpy(camera_sequence(1280, 720, 8, 300)).cuda()
    img = torch.empty((F, 720, 1280), dtype=torch.uint8, device="cuda")
    for i in range(0, F, 8):
        img[i:i + 8] = base
    images = {c: img for c in range(n_cams)}
    for _ in range(3):
        fe.step(images)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(5):
        fe.step(images)
    t_cpu = (time.perf_counter() - t0) / 5
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"n_cams {n_cams} chunk {chunk}: gpu {ms:.2f} ms/step = {ms * 1e3 / (n_cams * F):.2f} us/frame; host enqueue {t_cpu * 1e3:.2f} ms/step")
    fe.close()
