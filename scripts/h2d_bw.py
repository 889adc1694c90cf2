"""Pinned host-to-device bandwidth with and without binding the process to the GPU's NUMA node (gpurun_out/h2d_bw.log)."""
import os, subprocess, time
import torch
import pynvml


def bw(nbytes=256 << 20, reps=8):
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(1)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    up = nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
    e0.record()
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return up, nbytes * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


print(subprocess.run("lscpu | grep -i 'numa\\|^CPU(s)\\|Model name'; nvidia-smi topo -m | head -8", shell=True, capture_output=True, text=True).stdout)
torch.cuda.init()
print("affinity before:", len(os.sched_getaffinity(0)), "cpus")
print("H2D / D2H GB/s, default placement:", bw())
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(0).uuid)).encode())
pynvml.nvmlDeviceSetCpuAffinity(h)
print("affinity after :", len(os.sched_getaffinity(0)), "cpus", sorted(os.sched_getaffinity(0))[:4], "...")
print("H2D / D2H GB/s, bound to the GPU's node:", bw())
print("H2D / D2H GB/s, 32 MiB buffers:", bw(32 << 20, 32))
