"""Host<->device copy bandwidth per GPU, alone and together, with and without NUMA-local
pinned buffers.  Diagnostic for the e2e leg at N > 1.  Usage: python tests/tools/pcie_diag.py"""
import glob
import os
import subprocess
import sys
import threading
import time

import torch


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=60).stdout.strip()
    except Exception as e:  # noqa
        return f"<{e}>"


def gpu_numa(i):
    bus = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else None
    q = sh(f"nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i {i}").lower()
    q = q[4:] if q.startswith("0000") and len(q) > 12 else q
    for d in glob.glob("/sys/bus/pci/devices/*"):
        if d.lower().endswith(q[-12:]):
            return (open(d + "/numa_node").read().strip(), open(d + "/local_cpulist").read().strip())
    return (None, None)


def parse_cpulist(s):
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-"); out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def bw(devs, nbytes, bufs, reps=4):
    """simultaneous H2D + D2H on every device in devs; returns aggregate GB/s each way"""
    streams = {d: (torch.cuda.Stream(d), torch.cuda.Stream(d)) for d in devs}
    for d in devs:
        torch.cuda.synchronize(d)
    t0 = time.perf_counter()
    for _ in range(reps):
        for d in devs:
            hin, hout, din, dout = bufs[d]
            with torch.cuda.stream(streams[d][0]):
                din.copy_(hin, non_blocking=True)
            with torch.cuda.stream(streams[d][1]):
                hout.copy_(dout, non_blocking=True)
    for d in devs:
        torch.cuda.synchronize(d)
    dt = time.perf_counter() - t0
    return len(devs) * nbytes * reps / dt / 1e9


def main():
    n = torch.cuda.device_count()
    print("gpus", n, "cpus", os.cpu_count(), "affinity", sorted(os.sched_getaffinity(0)))
    print(sh("nvidia-smi topo -m"))
    print(sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"))
    print("numactl:", sh("which numactl"), "| nodes:", sh("ls /sys/devices/system/node | grep node"))
    print(sh("grep -E 'MemTotal|MemFree' /sys/devices/system/node/node*/meminfo"))
    nbytes = 1 << 30
    for mode in ("default", "numa-local"):
        bufs = {}
        base_aff = os.sched_getaffinity(0)
        for d in range(n):
            node, cpus = gpu_numa(d)
            if mode == "numa-local" and cpus:
                try:
                    os.sched_setaffinity(0, set(parse_cpulist(cpus)) & base_aff or base_aff)
                except OSError as e:
                    print("setaffinity failed", e)
            hin = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            hin.fill_(1)           # first touch happens in pin_memory(); fill keeps it honest
            hout = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
            hout.fill_(2)
            din = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{d}")
            dout = torch.zeros(nbytes, dtype=torch.uint8, device=f"cuda:{d}")
            bufs[d] = (hin, hout, din, dout)
            os.sched_setaffinity(0, base_aff)
            if mode == "default":
                print(f"gpu {d}: numa_node {node} local_cpulist {cpus}")
        for d in range(n):
            bw([d], nbytes, bufs, 1)
        print(mode, "each GPU alone (GB/s each way):", [round(bw([d], nbytes, bufs), 1) for d in range(n)])
        if n > 1:
            print(mode, "pairs:", {f"0+{d}": round(bw([0, d], nbytes, bufs), 1) for d in range(1, n)})
            print(mode, "all together, aggregate:", round(bw(list(range(n)), nbytes, bufs), 1))
        del bufs


if __name__ == "__main__":
    main()
