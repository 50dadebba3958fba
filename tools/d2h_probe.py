#!/usr/bin/env python
"""Bare concurrent device->host copy probe: what can this box's host side absorb, with no kernels at all?

    python tools/d2h_probe.py [--gpus 1,2,4,8] [--seconds 1.5] [--chunk-mb 78.8] [--mode pinned|wc|registered]
                              [--numa none|local|interleave] [--with-h2d]

For every N in --gpus it starts N processes (one per GPU); each calls b200s_copy_probe (include/b200_stereo.h): four
page-locked host buffers of --chunk-mb (the D2H payload of one C4 frame: rect L/R + float disparity + PointCloud2 =
78.8 MB), cudaMemcpyAsync device->host round-robin on two streams for --seconds, all processes released by one barrier.
Prints one JSON line per N: per-GPU GB/s and the aggregate.  bench.py's e2e.frac_of_copy_ceiling is measured against a
probe of the same kind inside the run; profiles/r02_d2h_probe.md holds the committed runs.
"""
import argparse
import ctypes
import json
import multiprocessing as mp
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODES = dict(pinned=0, wc=1, registered=2)


def set_mempolicy(mode, nodemask_bits):
    """set_mempolicy(2) through libc's syscall(): 0 default, 2 bind, 3 interleave (x86-64 syscall 238)."""
    libc = ctypes.CDLL(None, use_errno=True)
    mask = ctypes.c_ulong(nodemask_bits)
    return libc.syscall(238, ctypes.c_int(mode), ctypes.byref(mask), ctypes.c_ulong(64)) == 0


def numa_nodes():
    try:
        return sorted(int(d[4:]) for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except OSError:
        return [0]


def gpu_numa_node(index):
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("0000"):
            bus = bus[4:]
        return int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
    except Exception:
        return -1


def worker(rank, args, barrier, q):
    sys.path.insert(0, ROOT)
    from ros_gpu_stereo_processor_b200 import _capi
    lib = _capi.load()
    nodes = numa_nodes()
    node = gpu_numa_node(rank)
    note = "default"
    if args.numa == "interleave" and len(nodes) > 1:
        note = "interleave ok" if set_mempolicy(3, sum(1 << k for k in nodes)) else "interleave failed"
    elif args.numa == "local":
        note = ("bind node %d ok" % node if set_mempolicy(2, 1 << node) else "bind failed") if node >= 0 else "gpu numa node unknown"
    gbs = ctypes.c_double()
    # a very short first call creates the CUDA context before the barrier
    lib.b200s_copy_probe(rank, 1 << 20, 0.01, 0, 0, ctypes.byref(gbs))
    barrier.wait()
    rc = lib.b200s_copy_probe_ex(rank, int(args.chunk_mb * 1e6), args.seconds, MODES[args.mode], int(args.with_h2d), args.streams,
                                 args.buffers, ctypes.byref(gbs))
    q.put(dict(rank=rank, rc=rc, gbs=gbs.value, numa=note, gpu_numa_node=node))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", default="1")
    ap.add_argument("--seconds", type=float, default=1.5)
    ap.add_argument("--chunk-mb", type=float, default=78.8)
    ap.add_argument("--mode", default="pinned", choices=list(MODES))
    ap.add_argument("--numa", default="none", choices=["none", "local", "interleave"])
    ap.add_argument("--with-h2d", action="store_true", help="also stream a 4 MB host->device copy per D2H copy")
    ap.add_argument("--streams", type=int, default=2, help="copy streams per GPU")
    ap.add_argument("--buffers", type=int, default=4, help="page-locked host buffers per GPU")
    args = ap.parse_args()
    ctx = mp.get_context("spawn")
    try:
        avail = int(subprocess.run(["nvidia-smi", "--query-gpu=count", "--format=csv,noheader"], capture_output=True, text=True).stdout.split()[0])
    except Exception:
        avail = 1
    for n in [int(x) for x in args.gpus.split(",")]:
        if n > avail:
            print(json.dumps(dict(n_gpus=n, skipped="only %d GPUs visible" % avail)))
            continue
        barrier, q = ctx.Barrier(n), ctx.Queue()
        procs = [ctx.Process(target=worker, args=(r, args, barrier, q)) for r in range(n)]
        for p in procs:
            p.start()
        res = sorted([q.get(timeout=300) for _ in procs], key=lambda d: d["rank"])
        for p in procs:
            p.join(timeout=60)
        print(json.dumps(dict(n_gpus=n, mode=args.mode, numa=args.numa, with_h2d=args.with_h2d, chunk_mb=args.chunk_mb, streams=args.streams, buffers=args.buffers,
                              aggregate_gbs=sum(r["gbs"] for r in res), per_gpu_gbs=[round(r["gbs"], 2) for r in res],
                              all_ok=all(r["rc"] == 0 for r in res), numa_nodes=numa_nodes(),
                              gpu_numa_nodes=[r["gpu_numa_node"] for r in res], notes=sorted(set(r["numa"] for r in res)),
                              host_cpus=os.cpu_count())))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
