#!/usr/bin/env python
"""Bare concurrent device->host copy probe: what can this box's host side absorb, with no kernels at all?

    python tools/d2h_probe.py [--gpus 1,2,4,8] [--seconds 1.5] [--chunk-mb 78.8] [--mode pinned|registered|wc] [--numa none|local|interleave]

For every N in --gpus it starts N processes (one per GPU); each allocates 4 pinned host buffers of --chunk-mb (the
D2H payload of one C4 frame: rect L/R + float disparity + PointCloud2 = 78.8 MB) and streams cudaMemcpyAsync
device->host copies round-robin over them on two streams for --seconds, all processes released by one barrier.
Prints one JSON line per N: per-GPU GB/s and the aggregate.  bench.py's e2e.frac_of_copy_ceiling is measured against
the aggregate of the same N (profiles/r02_d2h_probe.md holds the committed runs).  torch only moves bytes here.
"""
import argparse
import ctypes
import json
import os
import sys
import time


def set_mempolicy(mode, nodemask_bits):
    """set_mempolicy(2) through libc's syscall(): 0 default, 2 bind, 3 interleave (x86-64 syscall 238)."""
    libc = ctypes.CDLL(None, use_errno=True)
    mask = ctypes.c_ulong(nodemask_bits)
    rc = libc.syscall(238, ctypes.c_int(mode), ctypes.byref(mask), ctypes.c_ulong(64))
    return rc == 0


def numa_nodes():
    try:
        return sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except OSError:
        return [0]


def gpu_numa_node(index):
    import subprocess
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("0000"):
            bus = bus[4:]
        return int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
    except Exception:
        return -1


def worker(rank, n, args, barrier, q):
    import torch
    torch.cuda.set_device(rank)
    nodes = numa_nodes()
    note = "default"
    if args.numa == "interleave" and len(nodes) > 1:
        note = "interleave ok" if set_mempolicy(3, sum(1 << k for k in nodes)) else "interleave failed"
    elif args.numa == "local":
        nd = gpu_numa_node(rank)
        if nd >= 0:
            note = ("bind node %d ok" % nd) if set_mempolicy(2, 1 << nd) else "bind failed"
        else:
            note = "gpu numa node unknown"
    nbytes = int(args.chunk_mb * 1e6)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev.fill_(7)
    if args.mode == "registered":
        arena = torch.empty(4 * nbytes, dtype=torch.uint8)
        arena.fill_(0)
        rc = torch.cuda.cudart().cudaHostRegister(arena.data_ptr(), arena.numel(), 0)
        assert int(rc) == 0, rc
        host = [arena[i * nbytes:(i + 1) * nbytes] for i in range(4)]
    else:
        host = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)]
        for h in host:
            h.fill_(0)
    hin = torch.empty(int(4.15e6), dtype=torch.uint8).pin_memory()
    din = torch.empty_like(hin, device="cuda")
    streams = [torch.cuda.Stream() for _ in range(2)]
    s_in = torch.cuda.Stream()
    for i in range(4):
        with torch.cuda.stream(streams[i & 1]):
            host[i].copy_(dev, non_blocking=True)
    torch.cuda.synchronize()
    barrier.wait()
    t0 = time.perf_counter()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in streams]
    ev0.record(streams[0])
    streams[1].wait_event(ev0)
    k = 0
    while time.perf_counter() - t0 < args.seconds:
        for i in range(4):
            with torch.cuda.stream(streams[i & 1]):
                host[i].copy_(dev, non_blocking=True)
            if args.with_h2d:
                with torch.cuda.stream(s_in):
                    din.copy_(hin, non_blocking=True)
            k += 1
        streams[0].synchronize()      # keep at most ~4 copies queued
    for s, e in zip(streams, ev1):
        e.record(s)
    torch.cuda.synchronize()
    ms = max(ev0.elapsed_time(e) for e in ev1)
    ok = bool((host[0][:: 4096] == 7).all().item())
    q.put(dict(rank=rank, gbs=k * nbytes / (ms * 1e-3) / 1e9, copies=k, ms=ms, ok=ok, numa=note, gpu_numa_node=gpu_numa_node(rank)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--chunk-mb", type=float, default=78.8)
    ap.add_argument("--mode", default="pinned", choices=["pinned", "registered"])
    ap.add_argument("--numa", default="none", choices=["none", "local", "interleave"])
    ap.add_argument("--with-h2d", action="store_true", help="also stream the 4.15 MB raw pair host->device per copy")
    args = ap.parse_args()
    import torch
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    avail = torch.cuda.device_count()
    for n in [int(x) for x in args.gpus.split(",")]:
        if n > avail:
            print(json.dumps(dict(n_gpus=n, skipped="only %d GPUs visible" % avail)))
            continue
        barrier, q = ctx.Barrier(n), ctx.Queue()
        procs = [ctx.Process(target=worker, args=(r, n, args, barrier, q)) for r in range(n)]
        for p in procs:
            p.start()
        res = sorted([q.get(timeout=120) for _ in procs], key=lambda d: d["rank"])
        for p in procs:
            p.join(timeout=60)
        print(json.dumps(dict(n_gpus=n, mode=args.mode, numa=args.numa, with_h2d=args.with_h2d, chunk_mb=args.chunk_mb,
                              aggregate_gbs=sum(r["gbs"] for r in res), per_gpu_gbs=[round(r["gbs"], 2) for r in res],
                              all_ok=all(r["ok"] for r in res), numa_nodes=numa_nodes(),
                              gpu_numa_nodes=[r["gpu_numa_node"] for r in res], notes=sorted(set(r["numa"] for r in res)),
                              host_cpus=os.cpu_count())))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
