"""Dev probe: the configs[1] extraction step (camera 1 + camera 2, 256 frames each) with the two extractors on ONE stream
(as bench.py runs them) against TWO streams (camera 2's kernels free to fill the SMs the latency-bound kernels of camera 1
leave idle), and against a split of every camera into two half-batches on two streams."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from multi_orb_slam_b200.extractor import ORBextractor
from multi_orb_slam_b200.synth import camera_sequence

W, H, F = 640, 480, 256
imgs = [torch.from_numpy(camera_sequence(W, H, 32, c)).cuda() for c in range(2)]
imgs = [torch.cat([i] * (F // i.shape[0]))[:F].contiguous() for i in imgs]


def run(tag, exs, streams, jobs):
    outs = []
    for (e, im), st in zip(jobs, streams):
        e.set_stream(st.cuda_stream)
        outs.append(e.extract_batch_device(im))
    def step():
        for (e, im), o in zip(jobs, outs):
            e.extract_batch_device(im, *o)
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < 0.5:
        step()
        torch.cuda.synchronize()
    K = 20
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for st in set(streams):
        st.wait_event(e0)
    for _ in range(K):
        step()
    for st in set(streams):
        ev = torch.cuda.Event(); ev.record(st); torch.cuda.current_stream().wait_event(ev)
    e1.record()
    torch.cuda.synchronize()
    print(f"{tag}: {e0.elapsed_time(e1) / K:.3f} ms per step, keypoints {sum(int(o[2].sum().item()) for o in outs)}")


sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
ex = [ORBextractor(nf, 1.2, 8, 20, 7, image_size=(W, H), max_batch=F, device=0) for nf in (1000, 500)]
run("one stream ", ex, [sa, sa], [(ex[0], imgs[0]), (ex[1], imgs[1])])
run("two streams", ex, [sa, sb], [(ex[0], imgs[0]), (ex[1], imgs[1])])
exh = [ORBextractor(nf, 1.2, 8, 20, 7, image_size=(W, H), max_batch=F // 2, device=0) for nf in (1000, 1000, 500, 500)]
halves = [(exh[0], imgs[0][:F // 2]), (exh[1], imgs[0][F // 2:]), (exh[2], imgs[1][:F // 2]), (exh[3], imgs[1][F // 2:])]
run("half batches, one stream ", exh, [sa, sa, sa, sa], halves)
run("half batches, two streams", exh, [sa, sb, sa, sb], halves)
