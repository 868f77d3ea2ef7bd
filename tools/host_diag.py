#!/usr/bin/env python
"""Host-side diagnosis for the multi-GPU e2e path: topology + concurrent pinned H2D/D2H bandwidth per rank."""
import os, sys, time, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
if rank == 0:
    print("nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    for cmd in (["nvidia-smi", "topo", "-m"], ["bash", "-c", "ls /sys/devices/system/node/ | grep node; cat /sys/devices/system/node/node*/cpulist; free -g | head -2"]):
        try:
            print(subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=20).stdout[:3000])
        except Exception as e:
            print("ERR", e)
p = torch.cuda.get_device_properties(lr)
bus = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
try:
    node = open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip()
except Exception as e:
    node = "?(%s)" % e
print("rank", rank, "gpu", lr, bus, "numa_node", node, "cpus", sorted(os.sched_getaffinity(0))[:4], "...", flush=True)
mode = sys.argv[1] if len(sys.argv) > 1 else "none"
if mode == "bind" and node not in ("?", "-1") and node.isdigit():
    cl = open("/sys/devices/system/node/node%s/cpulist" % node).read().strip()
    cpus = set()
    for part in cl.split(","):
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    os.sched_setaffinity(0, cpus)
    print("rank", rank, "bound to node", node, len(cpus), "cpus", flush=True)
hin = torch.empty(249_000_000, dtype=torch.uint8, pin_memory=True); hin.fill_(1)
hout = torch.empty(208_000_000, dtype=torch.uint8, pin_memory=True)
din = torch.empty_like(hin, device=dev); dout = torch.empty(208_000_000, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h, reps=10):
    if world > 1: dist.barrier()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1): din.copy_(hin, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): hout.copy_(dout, non_blocking=True)
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / reps
    if world > 1: dist.barrier()
    return t
for name, a, b in (("h2d", 1, 0), ("d2h", 0, 1), ("both", 1, 1)):
    run(a, b, 2); t = run(a, b)
    gb = (249e6 * a + 208e6 * b) / 1e9
    print("rank %d %s: %.2f ms  %.1f GB/s" % (rank, name, t * 1e3, gb / t), flush=True)
if world > 1: dist.destroy_process_group()
