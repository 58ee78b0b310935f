#!/usr/bin/env python
"""Runs the brute-force Hamming leg alone (configs[2] upper end) so that one `ncu --set full` capture of
k_bruteforce stays small:  ncu --set full --clock-control none --import-source on --kernel-name regex:k_bruteforce
-c 1 -f -o gpurun_out/prof_bf python tools/prof_bruteforce.py 65536"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from multi_orb_slam_b200.matcher import ORBmatcher
from multi_orb_slam_b200.synth import perturbed_descriptors, random_descriptors

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
m = ORBmatcher(0.9, True, device=0)
A = random_descriptors(n, 7)
B, _ = perturbed_descriptors(A, 8)
dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
idx, d1, d2 = (torch.empty((n,), dtype=torch.int32, device="cuda") for _ in range(3))
for _ in range(3):
    m.bruteforce_device(dB, dA, idx, d1, d2, th_dist=50, ratio=0.9)
m.sync()
print("accepted", int((idx >= 0).sum().item()))
# timing (CUDA events on the matcher's stream) when called with a second argument
if len(sys.argv) > 2:
    st = torch.cuda.Stream()
    m.set_stream(st.cuda_stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record(st)
    for _ in range(reps):
        m.bruteforce_device(dB, dA, idx, d1, d2, th_dist=50, ratio=0.9)
    e1.record(st)
    st.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{sys.argv[2]}: {ms:.3f} ms, {n * n / ms / 1e9:.1f} G pairs/s, frac of POPC peak {n * n / (ms * 1e-3) / (148 * 16 * 1.965e9 / 8):.3f}")
