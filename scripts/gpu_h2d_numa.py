"""Concurrent host->device bandwidth of every rank's pinned staging buffer, by the NUMA node the process is bound
to while it allocates (first touch) and copies.  torchrun --nproc-per-node N scripts/gpu_h2d_numa.py"""
import os, sys, glob, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1: dist.init_process_group("nccl", device_id=dev)
nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
def cpus_of(n):
    out = set()
    for part in open(f"/sys/devices/system/node/node{n}/cpulist").read().strip().split(","):
        lo, _, hi = part.partition("-"); out.update(range(int(lo), int(hi or lo) + 1))
    return out
all_cpus = os.sched_getaffinity(0)
pr = torch.cuda.get_device_properties(dev)
bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
try: sys_node = open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip()
except Exception as e: sys_node = f"? ({e})"
dst = torch.empty(256, 1000, 100, device=dev)
def measure(tag):
    src = torch.empty(256, 1000, 100).pin_memory(); src.fill_(1.0)
    for _ in range(2): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): dst.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    gbs = 10 * src.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    t = torch.tensor([gbs], device=dev); 
    if world > 1:
        g = [torch.empty_like(t) for _ in range(world)]; dist.all_gather(g, t)
        if rank == 0: print(tag, "GB/s per rank:", [round(float(x), 1) for x in g], "sum", round(sum(float(x) for x in g), 1), flush=True)
    else: print(tag, round(gbs, 1), flush=True)
    del src
print(f"rank {rank}: gpu {bdf} sysfs numa_node {sys_node}; nodes {nodes}; affinity {len(all_cpus)} cpus", flush=True)
measure("default (unbound)")
for n in nodes:
    c = cpus_of(n) & all_cpus
    if not c: continue
    os.sched_setaffinity(0, c)
    measure(f"all ranks bound to node {n}")
# alternate: rank r -> node r * len(nodes) // world
n = nodes[min(len(nodes) - 1, lr * len(nodes) // max(world, 1))]
c = cpus_of(n) & all_cpus
if c: os.sched_setaffinity(0, c)
measure("rank r bound to node r*nodes//world")
os.sched_setaffinity(0, all_cpus)
