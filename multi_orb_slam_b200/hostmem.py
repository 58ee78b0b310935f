"""Pinned host buffers close to the GPU that reads them.

On a multi-socket host the H2D copies of the streaming front end (pipeline.py) run at full PCIe rate only
when the pinned pages live on the NUMA node the GPU hangs off; with one process per GPU and no binding,
eight ranks were measured sharing 184 GB/s instead of 8 x 55 GB/s.  `numa_local(device)` binds the calling
thread to the GPU's CPUs (NVML's ideal affinity) while buffers are allocated — first touch places the pages —
and restores the previous affinity afterwards, so host-side workers keep every core."""
from __future__ import annotations

import os
from contextlib import contextmanager


def _nvml_handle(device: int):
    import pynvml
    import torch
    pynvml.nvmlInit()
    try:
        uuid = str(torch.cuda.get_device_properties(device).uuid)
        return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
    except Exception:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[device]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else device
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(idx)


@contextmanager
def numa_local(device: int):
    """Yields the number of CPUs the thread is bound to inside the block (None: binding unavailable)."""
    before = os.sched_getaffinity(0)
    bound = None
    try:
        nv, h = _nvml_handle(device)
        nv.nvmlDeviceSetCpuAffinity(h)
        now = os.sched_getaffinity(0)
        bound = len(now) if now else None
    except Exception:
        bound = None
    try:
        yield bound
    finally:
        try:
            os.sched_setaffinity(0, before)
        except OSError:
            pass
