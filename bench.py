#!/usr/bin/env python
"""Benchmark of the stereo hot path (BASELINE.json metric: stereo frames/s and Mdisp-evals/s; % INT-ALU roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C4] [--table C1,C2,C3,C5]
    torchrun --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU, frames sharded, no collective)

Headline workload (config.workload): BASELINE.json configs[3] = C4 -- 1920x1080 mono8 raw pair, camera_info rectification,
x-Sobel prefilter, StereoBM 256 disparities / block 11 / texture 10 / uniqueness 15, DisparityImage float payload and
PointCloud2 payload: the full rectify -> disparity -> pc2 chain.  One "step" = one pass of that chain over a batch of
FRAMES_PER_STEP frames per GPU (passes over the config's distinct synthetic frames).

  value  frames/s with the raw frames and every output resident in HBM (CUDA events on the slot streams)
  e2e    frames/s through the C ABI with pinned HOST buffers: H2D of the raw pair and D2H of the rectified pair,
         the float disparity and the PointCloud2 payload inside the timed region; frac_of_copy_ceiling relates it to a bare
         cudaMemcpyAsync D2H probe of the same bytes run by every rank at the same time (no kernels)
  roofline  dominant kernel (bm_vh_kernel): scalar-equivalent integer ops (7 per disparity evaluation,
         SURVEY.md 8(d)) over the CUDA-event duration of the matcher, against the INT peak measured on this GPU
         by the library's micro-benchmark: IADD3+IMAD interleaved = both integer pipes (frac), IADD3 alone
         (frac_vs_iadd3_peak) and BASELINE.md's theoretical 37.2 Tops/s (frac_vs_theoretical_37p2)
  roofline_other  rectify+prefilter and reproject+pack stages: algorithmic bytes over CUDA-event time against the HBM peak
  configs  the same measurements (short legs) for the other BASELINE configs C1, C2, C3, C5
  parity_checked  after every timed leg the outputs of one frame per slot are compared byte for byte with the OpenCV chain
  cpu_baseline  the reference's CPU path (cv::remap x2, cv::StereoBM, convertTo, reprojectImageTo3D, PointCloud2
         fill = ros_cpu_stereo_processing.launch) run with the real OpenCV (cv2) on this box's host cores
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: W, H, nd, block, rectify, speckle(win, range); batch = frames per launch (slot depth), slots = batches in flight
    # frames = distinct synthetic frames per GPU (enough for `slots` batches in flight)
    "C1": dict(W=752, H=480, nd=64, block=21, rectify=False, speckle=(0, 0), idx=1, batch=16, slots=4, frames=64),
    "C2": dict(W=1242, H=375, nd=128, block=15, rectify=False, speckle=(100, 4), idx=2, batch=16, slots=4, frames=64),
    "C3": dict(W=1280, H=720, nd=128, block=15, rectify=True, speckle=(0, 0), idx=3, batch=8, slots=4, frames=32),
    "C4": dict(W=1920, H=1080, nd=256, block=11, rectify=True, speckle=(0, 0), idx=4, batch=4, slots=4),    # batch 1: 1.5 % slower
    "C5": dict(W=3840, H=2160, nd=256, block=11, rectify=True, speckle=(0, 0), idx=5, batch=1, slots=4),
    # not a BASELINE config: C4 with the reference's default speckle filter on (GPU.cfg max_speckle_size 800,
    # max_speckle_diff 5 disparities = 80 raw units), i.e. what StereoProcessor::imageCb runs out of the box
    "C4s": dict(W=1920, H=1080, nd=256, block=11, rectify=True, speckle=(800, 80), idx=4, batch=1, slots=4),
    # not a BASELINE config either: C4 shape in the state the reference's matcher is in out of the box (GPU.cfg defaults
    # xsobel=False -> NORMALIZED_RESPONSE with the constructor's preFilterSize 5, uniqueness 0 and disp12MaxDiff 0 mirrored
    # from the cuda matcher's getters, src/GPUStereoProcessor.cpp:22-38, speckle filter 800 / 5 disparities)
    "C4r": dict(W=1920, H=1080, nd=256, block=11, rectify=True, speckle=(800, 80), idx=4, pft=0, ps=5, uniq=0, disp12=0, batch=1, slots=4),
}
FRAMES_PER_STEP = 64      # frames per GPU and step: passes over the config's distinct frames (16 unless the config says otherwise)
N_SLOTS = int(os.environ.get("B200S_BENCH_SLOTS", "0"))       # 0 = per config
BATCH = int(os.environ.get("B200S_BENCH_BATCH", "0"))         # 0 = per config
THEORETICAL_TOPS = 37.2   # 148 SMs x 128 int32 lanes x 1.965 GHz


def slots_of(c):
    return N_SLOTS or c.get("slots", 4)


def frames_of(c):
    return int(os.environ.get("B200S_BENCH_FRAMES", c.get("frames", 16)))


def passes_of(c):
    return max(1, FRAMES_PER_STEP // frames_of(c))


def batch_of(c):
    return max(1, min(BATCH or c.get("batch", 1), frames_of(c)))


def workload_name(c, name):
    return "%s: %dx%d mono8 raw pair, %s, %s cap31, StereoBM nd=%d block=%d tex10 uniq%d%s%s, DisparityImage f32 + PointCloud2" % (
        name, c["W"], c["H"], "rectify from camera_info" if c["rectify"] else "pre-rectified",
        "xsobel" if c.get("pft", 1) == 1 else "normalized-response ps%d" % c.get("ps", 9), c["nd"], c["block"], c.get("uniq", 15),
        (" disp12MaxDiff=%d" % c["disp12"]) if c.get("disp12", -1) >= 0 else "",
        (" speckle(%d,%d)" % c["speckle"]) if c["speckle"][0] else "")


def config_dict(c, name, frames_per_step):
    """The `config` object of the JSON line -- identical in both arms (ours / reference) for the same workload."""
    n = c["W"] * c["H"]
    return dict(workload=workload_name(c, name), frames_per_step_per_gpu=frames_per_step, distinct_frames_per_gpu=frames_of(c),
                sharding="independent frames per GPU, no collective",
                l2="no flush: each step cycles %d distinct frames; working set (inputs %.0f MB + products of the frames in flight) exceeds the 126 MB L2 for C3-C5"
                   % (frames_of(c), 2 * n * frames_of(c) / 1e6))


def evals_per_frame(c):
    W, H, nd, r = c["W"], c["H"], c["nd"], c["block"] // 2
    nominal = W * H * nd
    eff = max(W - (nd - 1) - 2 * r, 0) * max(H - 2 * r, 0) * nd
    return nominal, eff


def make_frames(c, n, seed0):
    from tools import synth   # input generation only (seeded synthetic pairs + scaled calibration)
    frames, cal = [], synth.scaled_calibration(c["W"], c["H"])
    for i in range(n):
        if c["rectify"]:
            L, R, _ = synth.synth_raw_pair(c["W"], c["H"], c["nd"], seed0 + i)
        else:
            L, R = synth.synth_pair(c["W"], c["H"], c["nd"], seed0 + i)
        frames.append((np.ascontiguousarray(L), np.ascontiguousarray(R)))
    return frames, cal


class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # under load = the upper half of the power samples (the sampler also sees set-up and the CPU legs)
        busy = [s for s, p in zip(sm, pw) if pw and p >= 0.5 * (min(pw) + max(pw))] or sm
        return dict(sm_mhz=float(np.median(busy)) if busy else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


def bind_to_gpu_numa(gpu_index):
    """Pins this process to the CPUs of the GPU's NUMA node (before any pinned allocation) so that the pinned host
    buffers of the end-to-end leg sit next to the GPU's PCIe root; returns a short description."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("0000"):
            bus = bus[4:]
        path = "/sys/bus/pci/devices/%s/" % bus
        node = open(path + "numa_node").read().strip()
        cpus = []
        for part in open(path + "local_cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return "numa node %s, %d cpus" % (node, len(allowed))
    except Exception as e:   # best effort only
        return "not bound (%s)" % type(e).__name__


def setup_processor(proc, c, cal):
    """Calibration + matcher state of config `c` on a GpuStereoProcessor (shared with tests/test_gpu_parity.py)."""
    W, H = c["W"], c["H"]
    info = lambda cc: dict(width=W, height=H, K=cc["K"], D=cc["D"], R=cc["R"], P=cc["P"])
    proc.initStereoModel(info(cal["left"]), info(cal["right"]))
    proc.setParams(numDisparities=c["nd"], blockSize=c["block"], minDisparity=0, preFilterType=c.get("pft", 1), preFilterSize=c.get("ps", 9),
                   preFilterCap=31, textureThreshold=10, uniquenessRatio=c.get("uniq", 15), speckleWindowSize=c["speckle"][0],
                   speckleRange=c["speckle"][1],
                   # -1 = BASELINE configs; environment override for experiments
                   disp12MaxDiff=int(os.environ.get("B200S_BENCH_DISP12", str(c.get("disp12", -1)))))
    if os.environ.get("B200S_BENCH_RECT_FLY"):
        proc.setRectifyOnTheFly(int(os.environ["B200S_BENCH_RECT_FLY"]) != 0)


def want_bits(c, capi):
    return capi.OUT_DISPARITY32F | capi.OUT_POINTCLOUD2 | (capi.OUT_RECT_L | capi.OUT_RECT_R if c["rectify"] else 0)


class CpuChain(object):
    """The reference's CPU path with the real OpenCV (cv2 = the library the reference calls): remap x2, StereoBM,
    convertTo, reprojectImageTo3D, PointCloud2 fill.  Checker / baseline only -- never on the product path."""

    def __init__(self, c, cal, threads):
        import cv2
        from oracle import oracle as O, cv2_ref as CV
        self.cv2, self.c = cv2, c
        cv2.setNumThreads(threads)
        p = O.BMParams(numDisparities=c["nd"], blockSize=c["block"], speckleWindowSize=c["speckle"][0], speckleRange=c["speckle"][1],
                       preFilterType=c.get("pft", 1), preFilterSize=c.get("ps", 9), uniquenessRatio=c.get("uniq", 15),
                       disp12MaxDiff=int(os.environ.get("B200S_BENCH_DISP12", str(c.get("disp12", -1)))))
        self.bm = CV.make_bm(p)
        W, H = c["W"], c["H"]
        self.maps = None
        if c["rectify"]:
            self.maps = [CV.rect_maps(cal[s]["K"], cal[s]["D"], cal[s]["R"], cal[s]["P"], W, H) for s in ("left", "right")]
        self.Q = O.stereo_Q(cal["left"]["P"], cal["right"]["P"])
        self.cxd = cal["left"]["P"][2] - cal["right"]["P"][2]

    def run(self, L, R):
        cv2, c = self.cv2, self.c
        H, W = c["H"], c["W"]
        out = {}
        if self.maps:
            L = cv2.remap(L, self.maps[0][0], self.maps[0][1], cv2.INTER_LINEAR)
            R = cv2.remap(R, self.maps[1][0], self.maps[1][1], cv2.INTER_LINEAR)
            out["rect_left"], out["rect_right"] = L, R
        d = self.bm.compute(L, R)
        df = (d.astype(np.float64) * (1.0 / 16.0) + (-self.cxd)).astype(np.float32)     # convertTo(CV_32F, 1/16, -(cx-cx'))
        xyz = cv2.reprojectImageTo3D(df, self.Q, handleMissingValues=True)
        # PointCloud2 fill (GpuSenderPc2.cpp:15-72), vectorised numpy instead of the reference's scalar loops
        pc = np.zeros((H, W, 8), np.float32)
        bad = (xyz[..., 2] == 10000.0) | np.isinf(xyz[..., 2])
        pc[..., :3] = np.where(bad[..., None], np.float32(np.nan), xyz)
        pcb = pc.view(np.uint8).reshape(H, W, 32)
        pcb[..., 16:19] = L[..., None]
        out["disparity32f"], out["pointcloud2"] = df, pcb
        return out

    def time(self, frames, budget_s):
        self.run(*frames[0])   # warm-up
        t0 = time.perf_counter()
        n = 0
        while True:
            self.run(*frames[n % len(frames)])
            n += 1
            if time.perf_counter() - t0 > budget_s and n >= 3:
                break
        return n / (time.perf_counter() - t0), n


def run_reference(args, c, name, rank, world):
    """--impl reference: the reference's CPU implementation of the path (cv2 = the OpenCV it calls), all host threads."""
    if rank != 0:
        return
    import cv2
    threads = os.cpu_count() or 1
    frames, cal = make_frames(c, 4, 1000 * c["idx"])
    chain = CpuChain(c, cal, threads)
    vals = []
    for s in range(args.warmup + args.steps):
        fps, n = chain.time(frames, budget_s=2.0)
        if s >= args.warmup:
            vals.append(fps)
    fps = float(np.mean(vals))
    nominal, eff = evals_per_frame(c)
    line = dict(impl="reference", metric="stereo_frames_per_sec", value=fps, unit="frames/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1000.0 / fps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="u8", data="synthetic", mdisp_evals_per_s=fps * nominal / 1e6,
                config=config_dict(c, name, FRAMES_PER_STEP),
                cpu_baseline=dict(value=fps, unit="frames/s", cores=threads, kind="reference",
                                  sample="cv2 %s (the OpenCV functions the reference calls: remap x2, StereoBM, convertTo, reprojectImageTo3D, "
                                         "PointCloud2 fill) on 4 distinct frames of the workload, ~2 s of frames per step" % cv2.__version__),
                e2e=dict(value=fps, unit="frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


class _DevMem(object):
    """Raw device pointer as a __cuda_array_interface__ object (torch only moves the bytes)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = dict(shape=(int(nbytes),), typestr="|u1", data=(int(ptr), False), version=3)


class ConfigRun(object):
    """One BASELINE config on one GPU: processor, synthetic frames, device-resident and pinned-host buffers, timed legs."""

    PRODUCTS = (("rect_left", "OUT_RECT_L", 1), ("rect_right", "OUT_RECT_R", 1), ("disparity32f", "OUT_DISPARITY32F", 4),
                ("pointcloud2", "OUT_POINTCLOUD2", 32))

    def __init__(self, name, c, rank, dev, dist):
        import torch
        import ros_gpu_stereo_processor_b200 as m
        from ros_gpu_stereo_processor_b200 import _capi as capi
        self.torch, self.capi, self.name, self.c, self.rank, self.dev, self.dist = torch, capi, name, c, rank, dev, dist
        self.W, self.H, self.n = c["W"], c["H"], c["W"] * c["H"]
        self.S, self.B = slots_of(c), batch_of(c)
        self.nframes = frames_of(c)
        self.frames, self.cal = make_frames(c, self.nframes, 1000 * c["idx"] + rank * self.nframes)
        self.proc = m.GpuStereoProcessor(dev)
        setup_processor(self.proc, c, self.cal)
        self.proc.configureSlots(self.S, self.H, self.W, self.B)
        self.want = want_bits(c, capi)
        self.products = [(k, getattr(capi, bit), es) for k, bit, es in self.PRODUCTS if self.want & getattr(capi, bit)]
        self.pins = []
        self.groups = [list(range(g, min(g + self.B, self.nframes))) for g in range(0, self.nframes, self.B)]
        self._dev_ready = self._host_ready = False

    # ---- device-resident leg: inputs in HBM (torch only owns the memory), products stay in the slot buffers ----
    def prepare_device(self):
        torch, capi = self.torch, self.capi
        self.dL = [torch.from_numpy(f[0]).cuda(self.dev) for f in self.frames]
        self.dR = [torch.from_numpy(f[1]).cuda(self.dev) for f in self.frames]
        torch.cuda.synchronize(self.dev)
        self.dev_batches = []
        for grp in self.groups:
            ios = (capi.FrameIO * len(grp))()
            for k in range(len(grp)):
                ios[k].want, ios[k].rectify, ios[k].inputs_on_device, ios[k].outputs_on_device = self.want, int(self.c["rectify"]), 1, 1
                ios[k].rows, ios[k].cols = self.H, self.W
            self.dev_batches.append(self.proc.makeBatch([self.dL[i].data_ptr() for i in grp], [self.dR[i].data_ptr() for i in grp], ios))
        self._dev_ready = True

    def step_device(self):
        for g, b in enumerate(self.dev_batches):
            self.proc.processBatchRaw(g % self.S, b)

    # ---- end-to-end leg: pinned host inputs and outputs, H2D + D2H inside the chain ----
    def pinned(self, nbytes, out=False):
        # B200S_BENCH_WC=1: output buffers write-combined (experiment; CPU reads of such memory are slow)
        a, ptr = self.proc.hostAlloc(nbytes, write_combined=out and os.environ.get("B200S_BENCH_WC") == "1")
        self.pins.append(ptr)
        return a, ptr

    def prepare_host(self):
        capi, n = self.capi, self.n
        hin = []
        for f in self.frames:
            a, pa = self.pinned(n); a[:] = f[0].ravel()
            b, pb = self.pinned(n); b[:] = f[1].ravel()
            hin.append((pa, pb))
        self.host_views = []       # [slot][frame in batch] -> {product: array}
        self.host_batches = []     # per group: batch args bound to the output buffers of slot g % S
        slot_ios = []
        for s in range(self.S):
            views, ios = [], (capi.FrameIO * self.B)()
            for k in range(self.B):
                v = {}
                ios[k].want, ios[k].rectify = self.want, int(self.c["rectify"])
                ios[k].rows, ios[k].cols = self.H, self.W
                for key, bit, es in self.products:
                    v[key], ptr = self.pinned(n * es, out=True)
                    setattr(ios[k], key, ptr)
                views.append(v)
            self.host_views.append(views)
            slot_ios.append(ios)
        for g, grp in enumerate(self.groups):
            ios = slot_ios[g % self.S]
            if len(grp) < self.B:
                ios = (capi.FrameIO * len(grp))(*[ios[k] for k in range(len(grp))])
            self.host_batches.append(self.proc.makeBatch([hin[i][0] for i in grp], [hin[i][1] for i in grp], ios))
        self.h2d_per_frame = 2 * n
        self.d2h_per_frame = sum(n * es for _, _, es in self.products)
        self._host_ready = True

    def step_host(self):
        for g, b in enumerate(self.host_batches):
            s = g % self.S
            self.proc.waitSlot(s)          # the slot's pinned output buffers are about to be overwritten
            self.proc.processBatchRaw(s, b)

    # ---- timing ----
    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def timed(self, step_fn, steps, warmup, passes):
        self.proc.syncParams()
        for _ in range(warmup):
            step_fn()
        self.barrier()
        l0 = self.proc.kernelLaunches()
        self.proc.batchBegin()
        for _ in range(steps * passes):
            step_fn()
        ms = self.proc.batchEnd()
        l1 = self.proc.kernelLaunches()
        self.barrier()
        # whole-job span = the slowest rank's device-timed span (MAX over ranks)
        from ros_gpu_stereo_processor_b200 import sharding
        _, ms = sharding.aggregate_throughput(1, ms, self.dist, "cuda:%d" % self.dev)
        return ms, l1 - l0

    # ---- parity self-check: the frames the last step left behind, one frame per slot, against the OpenCV chain ----
    def check(self, leg, chain):
        torch = self.torch
        frames = mismatches = 0
        bad = []
        last_group_of_slot = {}
        for g in range(len(self.groups)):
            last_group_of_slot[g % self.S] = g
        for s, g in sorted(last_group_of_slot.items()):
            self.proc.waitSlot(s)
            k = (g + s) % len(self.groups[g])          # a different position inside the batch per slot
            i = self.groups[g][k]
            want = chain.run(*self.frames[i])
            for key, bit, es in self.products:
                if leg == "device":
                    ptr, nbytes = self.proc.slotFrameDevicePtr(s, k, bit)
                    got = torch.as_tensor(_DevMem(ptr, nbytes), device="cuda:%d" % self.dev).cpu().numpy()
                else:
                    got = self.host_views[s][k][key]
                w = np.ascontiguousarray(want[key]).view(np.uint8).ravel()
                if not np.array_equal(got, w):
                    mismatches += 1
                    bad.append("%s/%s frame %d %s: %d bytes differ" % (self.name, leg, i, key, int((got != w).sum())))
            frames += 1
        return frames, mismatches, bad

    # ---- bare D2H probe on this GPU (no kernels): the copy ceiling of the end-to-end leg ----
    def copy_ceiling_fps(self, seconds=0.5):
        """b200s_copy_probe (the probe of tools/d2h_probe.py: cudaMemcpyAsync D2H of one frame's products, two streams,
        four page-locked buffers) run by every rank at the same time.  Returns (sum over the ranks, N x the slowest rank) in
        frames/s: the bench gives every rank the same number of frames, so the slowest PCIe path gates the job."""
        import ctypes as C
        gbs = C.c_double()
        self.barrier()
        rc = self.proc._lib.b200s_copy_probe(self.dev, int(self.d2h_per_frame), float(seconds), 0, 0, C.byref(gbs))
        fps = gbs.value * 1e9 / self.d2h_per_frame if rc == 0 else float("nan")
        total, slowest, n = fps, fps, 1
        if self.dist is not None:
            t = self.torch.tensor([fps], device="cuda:%d" % self.dev, dtype=self.torch.float64)
            tmin = t.clone()
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            self.dist.all_reduce(tmin, op=self.dist.ReduceOp.MIN)
            total, slowest, n = float(t.item()), float(tmin.item()), self.dist.get_world_size()
        return total, n * slowest

    def stage_times(self, reps=2):
        """Matcher alone + the other stages: one batch at a time on slot 0, CUDA events on its stream (graphs bypassed)."""
        self.proc.enableTiming(True)
        bm, stages = [], []
        for rep in range(reps):
            for g, b in enumerate(self.dev_batches):
                self.proc.processBatchRaw(0, b)
                self.proc.waitSlot(0)
                if rep == reps - 1:
                    t_ms, ev = self.proc.lastBmTime(0)
                    bm.append(t_ms / b[0])
                    st = self.proc.lastStageTimes(0)
                    stages.append({k: v / b[0] for k, v in st.items()})
        self.proc.enableTiming(False)
        return float(np.mean(bm)) * 1e-3, {k: float(np.mean([s[k] for s in stages])) * 1e-3 for k in stages[0]}

    def close(self):
        for ptr in self.pins:
            self.proc.hostFree(ptr)
        self.pins = []
        self.dL = self.dR = None
        self.proc.close()


def measure_config(name, c, args, rank, world, dev, dist, steps, warmup, passes, want_roofline, all_cpus):
    """Runs both timed legs (+ roofline extras on rank 0) of one config; returns the summary dict."""
    run = ConfigRun(name, c, rank, dev, dist)
    nominal, eff = evals_per_frame(c)
    out = dict(workload=workload_name(c, name), batch=run.B, slots=run.S)
    run.prepare_device()
    ms_dev, launches = run.timed(run.step_device, steps, warmup, passes)
    frames_total = run.nframes * passes * steps * world
    fps_dev = frames_total / (ms_dev * 1e-3)
    chain = None
    checks = dict(frames=0, mismatches=0, failures=[])
    if not args.no_check:
        chain = CpuChain(c, run.cal, len(all_cpus))
        f, mm, bad = run.check("device", chain)
        checks["frames"] += f; checks["mismatches"] += mm; checks["failures"] += bad
    run.prepare_host()
    ceiling_fps, ceiling_balanced_fps = run.copy_ceiling_fps()
    ms_e2e, _ = run.timed(run.step_host, steps, max(1, warmup), passes)
    fps_e2e = frames_total / (ms_e2e * 1e-3)
    if chain is not None:
        f, mm, bad = run.check("host", chain)
        checks["frames"] += f; checks["mismatches"] += mm; checks["failures"] += bad
    if dist is not None and not args.no_check:
        t = run.torch.tensor([checks["frames"], checks["mismatches"]], device="cuda:%d" % dev, dtype=run.torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        checks["frames"], checks["mismatches"] = int(t[0].item()), int(t[1].item())
    out.update(frames_per_s=fps_dev, e2e_frames_per_s=fps_e2e, ms_per_frame=ms_dev / (frames_total / world), launches_per_frame=launches / (frames_total / world),
               mdisp_evals_per_s=fps_dev * nominal / 1e6, mdisp_evals_effective_per_s=fps_dev * eff / 1e6,
               e2e=dict(value=fps_e2e, unit="frames/s", h2d_bytes_per_frame=run.h2d_per_frame, d2h_bytes_per_frame=run.d2h_per_frame,
                        copy_ceiling_frames_per_s=ceiling_fps, frac_of_copy_ceiling=fps_e2e / ceiling_fps,
                        copy_ceiling_equal_shares_frames_per_s=ceiling_balanced_fps, frac_of_equal_shares_ceiling=fps_e2e / ceiling_balanced_fps,
                        ceiling="b200s_copy_probe: bare cudaMemcpyAsync D2H of one frame's products, pinned, 2 streams / 4 buffers, no kernels, all ranks at once "
                                "(this run); equal_shares = N x the slowest rank's rate, what a job with the same number of frames on every rank can reach"),
               parity_checked=checks, _ms_dev=ms_dev, _ms_e2e=ms_e2e, _launches=launches)
    if want_roofline and rank == 0:
        t_bm, stages = run.stage_times()
        out.update(matcher_us=t_bm * 1e6, matcher_tevals_per_s=eff / t_bm / 1e12, _t_bm=t_bm, _stages=stages,
                   stage_us={k: v * 1e6 for k, v in stages.items()})
    out["_run"] = run
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = local_rank
    all_cpus = sorted(os.sched_getaffinity(0))
    numa = bind_to_gpu_numa(dev)
    name, c = args.config, CONFIGS[args.config]

    sampler = ClockSampler(dev)
    sampler.start()
    main = measure_config(name, c, args, rank, world, dev, dist, args.steps, args.warmup, passes_of(c), True, all_cpus)
    clocks = sampler.stop()
    run = main.pop("_run")
    nominal, eff = evals_per_frame(c)
    n = c["W"] * c["H"]

    roof = cpu_base = roof_other = None
    if rank == 0:
        t_bm, stages = main["_t_bm"], main["_stages"]
        names = ["iadd3", "vabsdiff4", "viadd16x2", "vimnmx16x2", "imad", "prmt", "lop3", "iadd3+imad"]
        int_peaks = {}
        for wch, nm in enumerate(names):
            ops, mhz = run.proc.intPeak(wch)
            int_peaks[nm] = ops / 1e12
        peak_tops = int_peaks["iadd3"]
        peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
        hbm_peak, hbm_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        if os.path.exists(peaks_file):
            hbm_peak, hbm_src = float(json.load(open(peaks_file))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        traffic, ncu_pipes = None, None
        tf = os.path.join(ROOT, "profiles", "bm_traffic.json")
        if name == "C4" and os.path.exists(tf):
            t_ = json.load(open(tf))
            traffic = t_["dram_bytes_read"] + t_["dram_bytes_write"]
            ncu_pipes = t_.get("pipes")
        achieved = eff * 7 / t_bm / 1e12
        alg_bytes = 2 * n + 2 * n     # two prefiltered u8 planes in, one s16 disparity plane out
        # roof: the best integer issue rate measured on this GPU -- IADD3 and IMAD interleaved, i.e. both half-rate integer
        # pipes busy (the kernel uses both); the single-pipe IADD3 rate SURVEY.md 8(d) suggests and BASELINE.md's
        # theoretical figure are reported beside it
        peak_mixed = max(int_peaks["iadd3+imad"], peak_tops)
        roof = dict(bound="int-alu", kernel="bm_vh_kernel<%d,%d> (warp-specialised SAD matcher, window sums in registers)" % (c["block"] // 2, c["nd"]),
                    achieved=achieved, peak=peak_mixed, unit="Tops/s (scalar-equivalent int32 lane-ops, 7 per disparity evaluation)",
                    frac=achieved / peak_mixed,
                    peak_source="measured: b200s_int_peak, IADD3 and IMAD dependent chains interleaved (ALU + IMAD pipes), all SMs, this run",
                    frac_vs_iadd3_peak=achieved / peak_tops, iadd3_peak=peak_tops,
                    frac_vs_theoretical_37p2=achieved / THEORETICAL_TOPS,
                    issue_slot_util=(ncu_pipes or {}).get("issue_active_pct"),
                    issue_slot_util_source="ncu capture committed under profiles/ (hardware counter; not measurable from CUDA events)",
                    note="packed instructions (VABSDIFF4, u16x2 adds/minima) execute 2-4 scalar-equivalent ops each, so the fraction of the "
                         "single-pipe IADD3 rate exceeds 1; ncu_pipes is the hardware view of the same kernel (profiles/, committed capture)",
                    ncu_pipes=ncu_pipes,
                    kernel_ms=t_bm * 1e3, evals_effective_per_launch=eff * run.B, frames_per_launch=run.B, gevals_per_s=eff / t_bm / 1e9, traffic=traffic,
                    per="kernel_ms, traffic and hbm.algorithmic_bytes are per frame: the launch time divided by frames_per_launch (CUDA events "
                        "on the matcher's stream), the DRAM bytes of a one-frame launch under ncu",
                    hbm=dict(achieved=alg_bytes / t_bm / 1e9, peak=hbm_peak, unit="GB/s", frac=alg_bytes / t_bm / 1e9 / hbm_peak,
                             algorithmic_bytes=alg_bytes, peak_source=hbm_src),
                    int_peaks_tops=int_peaks, share_of_step=t_bm * FRAMES_PER_STEP / (main["_ms_dev"] * 1e-3 / args.steps))
        # the kernels around the matcher: algorithmic bytes over the CUDA-event time of the stage
        rect_bytes = (2 * n + 2 * 4 * n + 4 * n) if c["rectify"] else (2 * n + 2 * n)    # raw pair + 4 B/px map pair -> rect + prefiltered pair
        pack_bytes = 2 * n + n + 32 * n                                                  # s16 disparity + grey -> 32 B records
        roof_other = {
            "rectify_prefilter": dict(kernel="rectify_xsobel_quad_kernel (both sides, one launch)", bound="hbm", algorithmic_bytes=rect_bytes,
                                      kernel_ms=stages["rectify_prefilter"] * 1e3, achieved=rect_bytes / stages["rectify_prefilter"] / 1e9,
                                      peak=hbm_peak, unit="GB/s", frac=rect_bytes / stages["rectify_prefilter"] / 1e9 / hbm_peak),
            "reproject_pack": dict(kernel="pack_lut_kernel (reproject + PointCloud2 records + float disparity; table path of reproject_pack_kernel)", bound="hbm", algorithmic_bytes=pack_bytes,
                                   kernel_ms=stages["reproject_pack"] * 1e3, achieved=pack_bytes / stages["reproject_pack"] / 1e9,
                                   peak=hbm_peak, unit="GB/s", frac=pack_bytes / stages["reproject_pack"] / 1e9 / hbm_peak),
            "to_float": dict(kernel="disparity_to_float_kernel", kernel_ms=stages["to_float"] * 1e3),
            "post": dict(kernel="validate / speckle kernels", kernel_ms=stages["post"] * 1e3),
            "peak_source": hbm_src, "timing": "CUDA events around each stage of slot 0, one frame at a time, this run"}

    # ---- the other BASELINE configs: short legs ----
    table = {}
    table[name] = {k: v for k, v in main.items() if not k.startswith("_")}
    for tname in [t for t in args.table.split(",") if t and t != name]:
        tc = CONFIGS[tname]
        tsteps = max(2, min(args.steps, 6))
        res = measure_config(tname, tc, args, rank, world, dev, dist, tsteps, 3, 1 if tname == "C5" else passes_of(tc), True, all_cpus)
        res.pop("_run").close()
        table[tname] = {k: v for k, v in res.items() if not k.startswith("_")}
        table[tname]["steps"] = tsteps
        if rank == 0 and roof is not None and "matcher_tevals_per_s" in res:
            table[tname]["frac"] = res["matcher_tevals_per_s"] * 7 / roof["peak"]

    if rank == 0 and world == 1 and not args.no_cpu:
        os.sched_setaffinity(0, all_cpus)     # the CPU baseline gets every host core again
        threads = len(all_cpus)
        import cv2
        fps_cpu, ncpu = CpuChain(c, run.cal, threads).time(run.frames, args.cpu_seconds)
        cpu_base = dict(value=fps_cpu, unit="frames/s", cores=threads, kind="reference",
                        sample="cv2 %s full chain (remap x2, StereoBM, convertTo, reprojectImageTo3D, PointCloud2 fill) on %d frames of the same workload, %d threads"
                               % (cv2.__version__, ncpu, threads))
    run.close()

    total_checks = dict(frames=sum(t["parity_checked"]["frames"] for t in table.values()),
                        mismatches=sum(t["parity_checked"]["mismatches"] for t in table.values()),
                        failures=sum((t["parity_checked"]["failures"] for t in table.values()), []),
                        against="cv2 chain (oracle/cv2_ref.py semantics), one frame per slot after each timed leg of each config, byte for byte")
    if rank == 0:
        if roof is not None:
            table[name]["frac"] = roof["frac"]
        e2e = dict(main["e2e"])
        e2e.update(h2d_bytes_per_step=e2e.pop("h2d_bytes_per_frame") * FRAMES_PER_STEP, d2h_bytes_per_step=e2e.pop("d2h_bytes_per_frame") * FRAMES_PER_STEP,
                   ms_per_step=main["_ms_e2e"] / args.steps, mdisp_evals_per_s=main["e2e_frames_per_s"] * nominal / 1e6)
        cfg = config_dict(c, name, FRAMES_PER_STEP)
        line = dict(metric="stereo_frames_per_sec", value=main["frames_per_s"], unit="frames/s", n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=main["_ms_dev"] / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="u8", data="synthetic",
                    mdisp_evals_per_s=main["mdisp_evals_per_s"], mdisp_evals_effective_per_s=main["mdisp_evals_effective_per_s"],
                    config=cfg, slots=run.S, frames_per_launch=run.B, e2e=e2e, gpu_launches=int(main["_launches"]), clocks=clocks, roofline=roof,
                    roofline_other=roof_other, cpu_baseline=cpu_base, parity_checked=total_checks, configs=table, host_binding=numa)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if total_checks["mismatches"]:
        sys.stderr.write("PARITY MISMATCH: %s\n" % "; ".join(total_checks["failures"][:8]))
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C4", choices=list(CONFIGS))
    ap.add_argument("--table", default="C1,C2,C3,C5", help="other configs measured with short legs into the `configs` table ('' = none)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-check", action="store_true", help="skip the parity self-check")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, CONFIGS[args.config], args.config, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
