#!/usr/bin/env python
"""H2D / D2H bandwidth of pinned host memory on this box, per NUMA placement of the pinning thread:
tells whether bench.py's e2e (31.5 GB in per step) sits at the link's ceiling."""
import os, time, glob
import torch

def numa_nodes():
    out = {}
    for d in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
        n = int(d.rsplit("node", 1)[1])
        cpus = open(os.path.join(d, "cpulist")).read().strip()
        s = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-"); s.update(range(int(a), int(b) + 1))
            elif part:
                s.add(int(part))
        out[n] = s
    return out

def gpu_node(idx=0):
    try:
        bus = torch.cuda.get_device_properties(idx).pci_bus_id if hasattr(torch.cuda.get_device_properties(idx), "pci_bus_id") else None
    except Exception:
        bus = None
    import subprocess
    try:
        q = subprocess.run(["nvidia-smi", "-i", str(idx), "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
        p = "/sys/bus/pci/devices/" + q.lower()[4:] + "/numa_node"
        return int(open(p).read()), q
    except Exception as e:
        return None, str(e)

def bw(nbytes, nstreams, reps=4, d2h=False):
    host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(nstreams)]
    dev = [torch.empty(nbytes, dtype=torch.uint8, device="cuda") for _ in range(nstreams)]
    st = [torch.cuda.Stream() for _ in range(nstreams)]
    for i in range(nstreams):
        with torch.cuda.stream(st[i]):
            dev[i].copy_(host[i], non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in range(nstreams):
            with torch.cuda.stream(st[i]):
                if d2h: host[i].copy_(dev[i], non_blocking=True)
                else: dev[i].copy_(host[i], non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    return reps * nstreams * nbytes / dt / 1e9

if __name__ == "__main__":
    torch.cuda.init()
    nodes = numa_nodes()
    gn, bus = gpu_node(0)
    print("numa nodes:", {k: len(v) for k, v in nodes.items()}, "gpu0 node:", gn, bus, "affinity now:", len(os.sched_getaffinity(0)), "cpus")
    all_cpus = os.sched_getaffinity(0)
    for n, cpus in nodes.items():
        use = cpus & all_cpus
        if not use:
            continue
        os.sched_setaffinity(0, use)
        for ns in (1, 2, 4):
            print(f"pinned by a thread on node {n}: H2D {ns} stream(s) x 315 MB: {bw(315_000_000, ns):.1f} GB/s")
        print(f"pinned by a thread on node {n}: D2H 1 stream: {bw(315_000_000, 1, d2h=True):.1f} GB/s")
    os.sched_setaffinity(0, all_cpus)
