"""What the host can take from N GPUs at once (run on a multi-GPU box): aggregate device-to-host bandwidth into pinned
buffers, all GPUs copying concurrently, with the buffers placed (a) by the main thread and (b) by a thread bound to the
CPUs of each GPU's NUMA node.  Prints the topology it found and one JSON line per experiment."""
import json
import os
import subprocess
import sys
import threading
import time

import torch

MB = 80  # per GPU per round, as one 1000-iteration shard of the cfg 2 count matrix


def numa_of_gpu(i):
    out = subprocess.run(["nvidia-smi", f"--id={i}", "--query-gpu=pci.bus_id", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    bus = out.lower()
    if len(bus.split(":")[0]) == 8:
        bus = bus[4:]
    try:
        return int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read()), bus
    except Exception:
        return -1, bus


def cpus_of_node(n):
    try:
        txt = open(f"/sys/devices/system/node/node{n}/cpulist").read().strip()
    except Exception:
        return set()
    out = set()
    for part in txt.split(","):
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out


def run(n_gpus, local_alloc, rounds=6):
    allowed = os.sched_getaffinity(0)
    bufs, srcs, streams = [None] * n_gpus, [], []
    for i in range(n_gpus):
        srcs.append(torch.full((MB << 18,), i, dtype=torch.int32, device=f"cuda:{i}"))
        streams.append(torch.cuda.Stream(device=i))

    def alloc(i):
        if local_alloc:
            cpus = cpus_of_node(numa_of_gpu(i)[0]) & allowed
            if cpus:
                os.sched_setaffinity(0, cpus)
        t = torch.empty((MB << 18,), dtype=torch.int32, pin_memory=True)
        t.zero_()
        bufs[i] = t

    th = [threading.Thread(target=alloc, args=(i,)) for i in range(n_gpus)]
    [t.start() for t in th]
    [t.join() for t in th]
    best = 1e9
    for r in range(rounds):
        for i in range(n_gpus):
            torch.cuda.synchronize(i)
        t0 = time.perf_counter()
        for i in range(n_gpus):
            with torch.cuda.stream(streams[i]):
                bufs[i].copy_(srcs[i], non_blocking=True)
        for i in range(n_gpus):
            streams[i].synchronize()
        best = min(best, time.perf_counter() - t0)
    print(json.dumps({"gpus": n_gpus, "buffers_placed": "per-GPU NUMA node" if local_alloc else "main thread", "MB_per_gpu": MB,
                      "ms": 1e3 * best, "aggregate_GBps": n_gpus * MB * 1.048576e-3 / best}), flush=True)


if __name__ == "__main__":
    n = torch.cuda.device_count()
    print(json.dumps({"gpus": n, "cpus_allowed": len(os.sched_getaffinity(0)), "gpu_numa": [numa_of_gpu(i) for i in range(n)],
                      "nodes": {k: len(cpus_of_node(k) & os.sched_getaffinity(0)) for k in range(8) if cpus_of_node(k)}}), flush=True)
    for k in (1, 2, 4, 8):
        if k <= n:
            run(k, False)
            run(k, True)
