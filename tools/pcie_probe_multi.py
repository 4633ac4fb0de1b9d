"""Aggregate host<->device ceiling of the box with N GPUs copying at the same time.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        tools/pcie_probe_multi.py [--numa]

Every rank copies 1 GiB page-locked buffers to / from its own GPU (H2D alone, D2H alone, both at once);
all ranks start together (gloo barrier), the aggregate is total bytes / slowest rank.  --numa pins each rank
to the CPUs of its GPU's NUMA node before it allocates (first touch decides where page-locked memory lives).
Rank 0 prints one JSON line.
"""
import argparse, json, os, sys, time
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--numa", action="store_true")
ap.add_argument("--mib", type=int, default=1024)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
numa = None
if a.numa:
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        cpus = open(f"/sys/devices/system/node/node{max(node, 0)}/cpulist").read().strip()
        ids = set()
        for part in cpus.split(","):
            lo, _, hi = part.partition("-")
            ids.update(range(int(lo), int(hi or lo) + 1))
        os.sched_setaffinity(0, ids)
        numa = {"node": node, "cpus": cpus}
    except Exception as e:  # noqa: BLE001
        numa = {"error": str(e)}
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("gloo")
n = a.mib << 20
h_a = torch.empty(n, dtype=torch.uint8).pin_memory()
h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
h_a.fill_(1), h_b.fill_(2)
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_a, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_b.copy_(d_b, non_blocking=True)


def both():
    h2d(), d2h()


res = {}
for name, fn, gb in (("h2d", h2d, 1), ("d2h", d2h, 1), ("both", both, 2)):
    fn(); torch.cuda.synchronize()
    best = None
    for _ in range(4):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t], dtype=torch.float64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        best = float(dt) if best is None else min(best, float(dt))
    res[name] = {"aggregate_GBs": world * n * gb / best / 1e9, "per_gpu_GBs": n * gb / best / 1e9, "ms": best * 1e3}
if rank == 0:
    print(json.dumps({"n_gpus": world, "mib_per_gpu": a.mib, "numa_pinning": numa, **res}), flush=True)
if world > 1:
    dist.destroy_process_group()
