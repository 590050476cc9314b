"""GPU-box probe: pinned H2D bandwidth from each NUMA node's CPUs (where does cudaHostAlloc land?)."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from nisqa_b200 import dist as D
pr = torch.cuda.get_device_properties(0)
print("gpu0 pci", getattr(pr, "pci_domain_id", None), pr.pci_bus_id, pr.pci_device_id, "cpus allowed", len(os.sched_getaffinity(0)))
nodes = sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
print("numa nodes", nodes)
allc = os.sched_getaffinity(0)
dev = torch.empty(61_440_000, dtype=torch.uint8, device="cuda")
def bw(tag):
    host = torch.empty(61_440_000, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    torch.cuda.synchronize()
    for _ in range(3): dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20): dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    print("%-28s %.1f GB/s" % (tag, 61.44e6 / dt / 1e9), flush=True)
bw("default placement")
for n in nodes:
    try:
        spec = open("/sys/devices/system/node/node%d/cpulist" % n).read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus.update(range(int(a), int(b) + 1))
            elif part: cpus.add(int(part))
        cpus &= allc
        if not cpus: continue
        os.sched_setaffinity(0, cpus)
        bw("bound to node %d (%d cpus)" % (n, len(cpus)))
    except Exception as ex:
        print("node", n, "failed", ex)
os.sched_setaffinity(0, allc)
print("bind_to_gpu_numa ->", D.bind_to_gpu_numa(0))
