#!/usr/bin/env python
"""Benchmark of the rain-rendering hot path (BASELINE.json metric: rainy frames/s at 1242x375,
25 mm/h, HBM GB/s vs roofline).

  python bench.py --gpus N --steps K --warmup W            # the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A step = one pass of the hot path over one batch of 64 synthetic KITTI-sized frames
(BASELINE config C2).  ``value`` is measured with the batch already resident in HBM,
``e2e`` through rr_render_frames with pinned HOST buffers (H2D and D2H inside the timed region).
One rank per GPU (torchrun), frames sharded, no data-path collective; the only collective is the
NCCL broadcast of the streak DB at init.  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rain_rendering_b200 import synth  # noqa: E402
from rain_rendering_b200.streaks import STREAK_DTYPE  # noqa: E402

WORKLOAD = "C2"
BATCH = 64
METRIC = "rainy frames/sec at 1242x375, 25mm/hr"
REC_BYTES = STREAK_DTYPE.itemsize


def algorithmic_bytes_per_frame(W, H, n_streaks):
    """compulsory traffic at the C-ABI boundary of one frame of the device-resident arm: uint8 BGR in,
    uint16 depth samples in (round 1: float32, 4 bytes), float32 BGR + float32 mask + uint8 BGR + uint8 mask index out,
    streak records in."""
    return 3 * W * H + 2 * W * H + (12 + 4 + 3 + 1) * W * H + REC_BYTES * n_streaks


def build_batch(wl, rank, batch):
    """Synthetic batch of one rank: frames + simulator frames (host arrays).  Frame i of rank r is synth frame
    1000 * r + i, rendered with simulator frame i of particle set 1000 + r and np.random.seed(i)."""
    from rain_rendering_b200 import api, streaks as S
    W, H = wl["W"], wl["H"]
    cam = synth.CAMERAS[wl["dataset"]]
    db = synth.make_streak_db(0)
    frames = [synth.make_frame(W, H, 1000 * rank + i) for i in range(batch)]
    bgr = np.stack([f[0] for f in frames])
    depth = np.stack([f[1] for f in frames])
    d16 = np.rint(depth * 256.0).astype(np.uint16)           # the depth PNG's samples (synth depth is a multiple of 1/256 m)
    assert np.array_equal(d16.astype(np.float32) / 256., depth)
    parts = synth.make_particles(W, H, batch, wl["n_xml"], cam["cam_exposure"], seed=1000 + rank)
    with tempfile.TemporaryDirectory() as d:
        xml = os.path.join(d, "sim_camera0.xml")
        synth.write_particles_xml(parts, xml, cam["cam_exposure"])
        sim = S.load_streaks_from_xml(xml, 1, W, H)
    recs, offs = api.assemble_batch(sim[:batch], list(range(batch)), W, H, db.ratios)
    return db, bgr, depth, d16, sim[:batch], recs.copy(), offs


class ClockSampler:
    """SM clocks and throttle reasons sampled through NVML on a thread DURING the timed region."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = None
        self._thr = None

    def start(self):
        import threading
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[self.gpu].isdigit() else self.gpu
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            return
        names = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        self._stop = threading.Event()

        def loop():
            while not self._stop.is_set():
                try:
                    self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for n, bit in names.items():
                        if r & bit:
                            self.reasons.add(n)
                except Exception:
                    pass
                self._stop.wait(0.02)

        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join(timeout=2)
        out = dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(self.sm))
        if self.sm:
            out["sm_mhz"] = float(np.median(self.sm))
        return out


# ----------------------------------------------------------------------------------------------
# CPU baseline: the oracle port on the host cores (test infrastructure used as the checker/baseline)
# ----------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(wl_name, base_seed):
    """Pool initializer: everything that is not the per-frame path (imports, synthetic inputs, the
    per-camera tables) is prepared once per worker, outside the timed region.  Worker k renders frame
    base_seed + k OF THE GPU ARM'S BATCH (rank 0): the same image, depth, simulator frame and RNG seed."""
    import multiprocessing as mp
    import cv2
    cv2.setNumThreads(1)
    from oracle import rain_oracle as ro
    ident = mp.current_process()._identity
    k = (ident[0] - 1) if ident else 0
    wl = synth.WORKLOADS[wl_name]
    W, H = wl["W"], wl["H"]
    cam_s = synth.CAMERAS[wl["dataset"]]
    cam = ro.Camera(W=W, H=H, focal_mm=cam_s["cam_focal"], f_number=cam_s["cam_f_number"], exposure_ms=cam_s["cam_exposure"],
                    gain=cam_s["cam_gain"], fallrate=wl["fallrate"])
    db = synth.make_streak_db(0)
    frame_idx = (base_seed + k) % BATCH
    bgr, depth = synth.make_frame(W, H, frame_idx)                       # rank 0: synth frame 1000 * 0 + i
    parts = synth.make_particles(W, H, BATCH, wl["n_xml"], cam_s["cam_exposure"], seed=1000)
    with tempfile.TemporaryDirectory() as d:
        xml = os.path.join(d, "sim_camera0.xml")
        synth.write_particles_xml(parts[frame_idx:frame_idx + 1], xml, cam_s["cam_exposure"])
        streaks = ro.load_streaks_from_xml(xml, 1, W, H)[0]
    tables = ro.build_env_tables(W, H, cam.focal_m)
    omega = ro.solid_angles(H, tables.W_env)
    _W.update(ro=ro, cam=cam, db=db, bgr=bgr, depth=depth, streaks=streaks, tables=tables, omega=omega, frame_idx=frame_idx)


def _cpu_worker(step):
    import copy
    w = _W
    streaks = copy.deepcopy(w["streaks"])       # render_frame mutates the end points (wind write-back), like the reference
    t0 = time.perf_counter()
    r = w["ro"].render_frame(w["bgr"], w["depth"], streaks, w["db"].textures, w["db"].ratios, w["cam"], w["frame_idx"],
                             w["tables"], w["omega"], f32_mode="native")
    return time.perf_counter() - t0, r.n_streaks


def cpu_steps(steps, warmup, n_workers, budget_s=None):
    """Each step: n_workers processes render one C2 frame each (the reference's own scale-out is
    process-level over disjoint frames, main_threaded.py:109-176).  Returns per-step wall seconds.
    A frame takes several seconds on a core, so a step cannot be made shorter than that; with ``budget_s`` the run
    stops early (after at least one timed step) once warm-up + timed steps have used the budget, so that the arm
    always ends within a few minutes whatever --steps says."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    times, streaks = [], []
    with ctx.Pool(n_workers, initializer=_cpu_init, initargs=(WORKLOAD, 0)) as pool:
        pool.map(_cpu_ready, range(n_workers * 4))          # all workers initialised before anything is timed
        t_start = time.perf_counter()
        for s in range(warmup + steps):
            t0 = time.perf_counter()
            res = pool.map(_cpu_worker, [s] * n_workers, chunksize=1)
            dt = time.perf_counter() - t0
            if s >= warmup:
                times.append(dt)
                streaks += [r[1] for r in res]
            used = time.perf_counter() - t_start
            if budget_s is not None and used + dt > budget_s:
                if s < warmup:
                    warmup = s + 1                            # no time for more warm-up: the next step is timed
                elif times:
                    break
    return times, streaks


def _cpu_ready(i):
    time.sleep(0.05)
    return bool(_W)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_cores()
    # one process per frame like main_threaded.py (at most 10 children there).  More than ~16 concurrent renders
    # do not raise the throughput on the B200 host (memory-bound masked gathers: 64 processes 1.53 frames/s,
    # 32 processes 1.8 frames/s on 128 cores) and only stretch the step, so the sample is capped at 16.
    workers = max(1, min(cores, 16))
    wl = synth.WORKLOADS[WORKLOAD]
    budget = float(os.environ.get("RR_REFERENCE_BUDGET_S", "240"))
    times, streaks = cpu_steps(args.steps, args.warmup, workers, budget_s=budget)
    total = float(np.sum(times))
    fps = workers * len(times) / total
    sample = "%d processes x 1 frame per step (frames 0..%d of the GPU arm's rank-0 batch: %dx%d, %d mm/h, ~%d streaks/frame), oracle port of the " \
             "reference algorithm, cv2 threads 1 per process, per-camera tables/solid angles and synthetic inputs prepared outside the timed " \
             "region; host has %d cores" % (workers, workers - 1, wl["W"], wl["H"], wl["fallrate"], int(np.mean(streaks)), cores)
    if len(times) < args.steps:
        sample += "; %d of the %d requested steps timed (a step is %.1f s of CPU work per core; wall-clock budget %d s, RR_REFERENCE_BUDGET_S)" % (
            len(times), args.steps, total / len(times), int(budget))
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": len(times), "steps_requested": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2 KITTI 1242x375 25mm/hr", "frames_per_step": workers, "streaks_per_frame": float(np.mean(streaks))},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": workers, "host_cores": cores, "kind": "port", "sample": sample},
            "host_cores": cores,
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# the CUDA path
# ----------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from rain_rendering_b200 import api, dist as rdist
    numa = rdist.bind_to_gpu_numa(rdist.env_rank()[2])         # before any page-locked buffer exists
    rank, world, local = rdist.init_process_group("nccl")
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node %d" % args.gpus
    torch.cuda.set_device(local)
    wl = synth.WORKLOADS[WORKLOAD]
    W, H = wl["W"], wl["H"]
    cam = synth.CAMERAS[wl["dataset"]]
    batch = args.batch
    db, bgr, depth, d16, sim, recs, offs = build_batch(wl, rank, batch)
    n_streaks = int(offs[-1])
    # `lanes` independent contexts on this GPU take the steps in turn (api.RainLanes): the kernels of consecutive batches
    # overlap, in the device-resident arm and end to end alike
    n_lanes = max(1, args.lanes)
    lanes = api.RainLanes(local, n_lanes)
    for c in lanes.ctxs:
        rdist.broadcast_streak_db(c, db.textures if rank == 0 else None, src=0)   # the one collective (once per lane, at set-up)
    lanes.set_camera(W, H, cam["cam_focal"], cam["cam_f_number"], cam["cam_exposure"], cam["cam_gain"], wl["fallrate"], 1.0, batch)
    ctx = lanes.ctxs[0]
    from rain_rendering_b200 import _lib
    lib, C = ctx.lib, __import__("ctypes")
    streams = []
    for c in lanes.ctxs:
        stream_ptr = C.c_void_p()
        lib.rr_stream(c.h, C.byref(stream_ptr))
        streams.append(torch.cuda.ExternalStream(stream_ptr.value, device=torch.device("cuda", local)))
    stream = streams[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm ----------------------------------------------------------------
    # inputs as the files hold them (uint8 image, uint16 depth samples), every output form the library offers
    dev = torch.device("cuda", local)
    d_bgr = torch.from_numpy(bgr).to(dev)
    d_depth = torch.from_numpy(d16.view(np.int16)).to(dev)
    d_recs = torch.from_numpy(recs.view(np.uint8).reshape(-1)).to(dev)
    offs_c = np.ascontiguousarray(offs)
    outs, io_dev = [], []
    for _ in range(n_lanes):                # the inputs are shared (read-only), every lane writes its own outputs
        o = dict(bgr=torch.empty((batch, H, W, 3), dtype=torch.float32, device=dev), mask=torch.empty((batch, H, W), dtype=torch.float32, device=dev),
                 u8=torch.empty((batch, H, W, 3), dtype=torch.uint8, device=dev), idx8=torch.empty((batch, H, W), dtype=torch.uint8, device=dev))
        outs.append(o)
        io_dev.append(_lib.FrameIO(d_bgr.data_ptr(), d_depth.data_ptr(), _lib.DEPTH_U16_256, 0, d_recs.data_ptr(), _lib.ptr(offs_c).value,
                                   o["bgr"].data_ptr(), o["mask"].data_ptr(), o["u8"].data_ptr(), o["idx8"].data_ptr(), None, None))
    d_out_u8, d_out_idx8 = outs[0]["u8"], outs[0]["idx8"]

    def step_device(i, sync=0):
        k = i % n_lanes
        _lib.check(lib.rr_render_frames_device_io(lanes.ctxs[k].h, batch, C.byref(io_dev[k]), sync), "rr_render_frames_device_io")

    def join_lanes():                       # lane 0's stream waits for the work queued on the others
        for st in streams[1:]:
            ev = torch.cuda.Event()
            ev.record(st)
            stream.wait_event(ev)

    for i in range(max(args.warmup, 3) * n_lanes):
        step_device(i, 1)                   # synchronous: the first one sizes the lane's patch arena
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = lanes.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t_enq = time.perf_counter()
    for i in range(args.steps):
        step_device(i, 0)                   # queued: a lane's steps follow each other on its stream, the lanes overlap
    t_enq = time.perf_counter() - t_enq     # host time to enqueue the steps (the GPU is still working on them)
    join_lanes()
    e1.record(stream)
    barrier()
    ms_dev = e0.elapsed_time(e1)
    # per-stage CUDA-event durations of every lane's LAST step of the timed region (rr_synchronize settles them)
    stage_ms = {k: 0.0 for k in ctx.timings()}
    used = [c for j, c in enumerate(lanes.ctxs) if j < args.steps]
    for c in used:
        c.synchronize()
        for k, v in c.timings().items():
            stage_ms[k] += v * args.steps / len(used)          # scaled to the sum over the steps the lines below divide by
    launches = lanes.kernel_launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # every stage's own duration: the same step a few more times with the streak chain on the main stream (in the timed region
    # above it runs beside the frame chain, so the stage times there overlap and add up to more than the step)
    ctx.set_option("serial", 1)
    solo_ms = {k: 0.0 for k in ctx.timings()}
    for _ in range(3):
        step_device(0, 1)
        for k, v in ctx.timings().items():
            solo_ms[k] += v / 3.0
    ctx.set_option("serial", 0)
    # ---- end-to-end arm: pinned HOST buffers through the public API ---------------------------------
    # rr_submit_frames_io / rr_wait_frames with two sets of host buffers: while batch k renders, batch k+1 is
    # copied in and batch k-1 is copied out.  Every step builds its streak records from the simulator frames
    # (in-frame filter + the NumPy RNG mirror, rr_host_assemble_batch -- host work a caller pays per frame), copies its
    # inputs host->device in the forms the files hold (uint8 image, uint16 depth samples) and copies back what
    # Generator.run saves (generator.py:466-467): the uint8 image and the mask as plt.imsave's colormap index + range.
    n_sets = lanes.capacity                  # a host set is reused when the submission that used it has been waited for
    sets = []
    for _ in range(n_sets):
        hb = dict(bgr=api.PinnedBuffer(bgr.shape, np.uint8), depth=api.PinnedBuffer(d16.shape, np.uint16),
                  idx8=api.PinnedBuffer((batch, H, W), np.uint8),
                  u8=api.PinnedBuffer((batch, H, W, 3), np.uint8), rng=api.PinnedBuffer((batch, 2), np.float64))
        hb["bgr"].array[...] = bgr; hb["depth"].array[...] = d16
        sets.append(hb)
    # the records rotate over one more buffer: batch i is assembled while the oldest submission may still be copying out
    rec_sets = [api.PinnedBuffer((sum(len(f) for f in sim),), STREAK_DTYPE) for _ in range(n_sets + 1)]
    seeds = list(range(batch))

    def assemble(i):
        return api.assemble_batch(sim, seeds, W, H, db.ratios, out=rec_sets[i % (n_sets + 1)].array)

    def submit(i, r, o):
        hb = sets[i % n_sets]
        lanes.submit_frames(hb["bgr"].array, hb["depth"].array, r, o, None, None, hb["u8"].array, hb["idx8"].array, None, hb["rng"].array)

    def run_e2e(n):
        # host work of the next batch first, then the wait for the oldest one: the wait -> assemble -> copy-in chain would
        # otherwise be as long as the GPU's step
        for i in range(n):
            r, o = assemble(i)
            if lanes.inflight == lanes.capacity:
                lanes.wait_frames()
            submit(i, r, o)
        while lanes.inflight:
            lanes.wait_frames()

    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    link = None
    if args.skip_e2e:
        f0.record(stream); f1.record(stream)
    else:
        for c in lanes.ctxs:
            c.render_frames(bgr, d16, recs, offs_c, want=("u8", "idx8"))    # synchronous call: sizes the lane's host-path buffers
        run_e2e(max(args.warmup, 3) * n_lanes)
        barrier()
        f0.record(stream)
        run_e2e(args.steps)                 # returns when every output is in host memory; lane 0's stream is idle by then
        f1.record(stream)
    barrier()
    ms_e2e = max(f0.elapsed_time(f1), 1e-6)
    h2d = int(bgr.nbytes + d16.nbytes + recs.nbytes + offs_c.nbytes)
    d2h = int(sets[0]["idx8"].array.nbytes + sets[0]["u8"].array.nbytes + sets[0]["rng"].array.nbytes)
    checksum = float(sets[0]["idx8"].array.sum(dtype=np.int64))
    same = bool(np.array_equal(sets[0]["u8"].array, d_out_u8.cpu().numpy()) and np.array_equal(sets[0]["idx8"].array, d_out_idx8.cpu().numpy())) \
        if not args.skip_e2e else None
    if not args.skip_e2e:
        # the host link's ceiling, measured now with every rank copying at once (tools/pcie_ceiling.py is the stand-alone form)
        barrier()
        link = api.host_link_probe(local, 256, 0.6)
        barrier()
    # ---- max over ranks ------------------------------------------------------------------------
    t = torch.tensor([ms_dev, ms_e2e], dtype=torch.float64, device=dev)
    lk = torch.tensor(list(link or (0.0, 0.0)), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lk, op=dist.ReduceOp.SUM)
    ms_dev, ms_e2e = float(t[0]), float(t[1])
    link_total = float(lk[0] + lk[1])
    total_frames = batch * world * args.steps
    if rank == 0:
        value = total_frames / (ms_dev / 1000.0)
        e2e = total_frames / (ms_e2e / 1000.0)
        n_per_frame = n_streaks / batch
        balg = algorithmic_bytes_per_frame(W, H, n_per_frame)
        # dominant kernel (stage) of the device-resident step, timed live with CUDA events on the library's stream.  Inside the
        # timed region the stages of the two chains and of the lanes overlap, so an event span there is not a kernel's duration:
        # the roofline uses the stage's own duration (the pass with one kernel at a time right after the timed region, which is
        # also what the ncu launch list shows); the in-step spans are reported beside it.
        kern = {k: v / args.steps for k, v in stage_ms.items() if k not in ("h2d", "d2h", "total")}
        solo = {k: v for k, v in solo_ms.items() if k not in ("h2d", "d2h", "total")}
        dom = max(solo, key=solo.get)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = balg * batch / (solo[dom] / 1000.0) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get(dom)
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": "C2 KITTI 1242x375 25mm/hr", "batch_frames_per_gpu": batch, "streaks_per_frame": n_per_frame,
                           "lanes_per_gpu": n_lanes,
                           "lanes": "independent contexts on one GPU taking the steps (64-frame batches) in turn, so that consecutive batches overlap "
                                    "(api.RainLanes); --lanes 1 = one context, one step at a time",
                           "device_arm": "uint8 image + uint16 depth resident in HBM -> float32 image, float32 mask, uint8 image, uint8 mask index",
                           "l2": "inputs larger than L2 (%.0f MB per step per GPU, no flush)" % ((h2d) / 1e6),
                           "parallelism": "frames sharded x%d, no data-path collective" % world},
                "clocks": clocks, "gpu_launches": launches,
                "host_enqueue_ms_per_step": 1000.0 * t_enq / args.steps,
                "host_cores": host_cores(),
                "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps,
                        "checksum_mask": checksum, "equals_device_arm": same, "numa_bind": numa,
                        "host_link_gbs": link_total if link else None,
                        "host_link_note": "sum over ranks of concurrent page-locked H2D + D2H copies (rr_host_link_probe), measured in this run right after the timed region",
                        "link_frac": ((h2d + d2h) * world * args.steps / (ms_e2e / 1000.0) / 1e9 / link_total) if link and link_total > 0 else None,
                        "inputs": "uint8 BGR image + uint16 depth samples (what the PNG files hold) + streak records built per step from the simulator frames",
                        "outputs": "uint8 BGR image + rain mask as plt.imsave's colormap index (uint8) and its (min, max) -- what Generator.run saves",
                        "api": "rr_host_assemble_batch + rr_submit_frames_io / rr_wait_frames round-robin over %d contexts (api.RainLanes), "
                               "%d host buffer sets" % (n_lanes, lanes.capacity)},
                "stage_ms_solo": {k: v for k, v in solo_ms.items() if k not in ("h2d", "d2h", "total")},
                "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "kernel_ms": solo[dom], "stage_span_ms_in_step": kern[dom],
                             "frac_in_step": (balg * batch / (kern[dom] / 1000.0) / 1e9) / peak if kern[dom] > 0 else None,
                             "note": "achieved / frac: algorithmic bytes of one 64-frame launch / the dominant stage's own duration (CUDA events, one "
                                     "kernel at a time, measured in this run right after the timed region; agrees with the ncu launch list).  "
                                     "stage_ms / *_in_step: event spans of every lane's last step inside the timed region, where the streak chain "
                                     "(raster, blur) runs beside the frame chain (fog, env, setup) and the lanes run beside each other: the spans "
                                     "overlap and add up to more than the step",
                             "traffic": traffic, "peak_source": "measured" if "hbm_gbs" in peaks else "fallback",
                             "algorithmic_bytes_per_frame": balg, "whole_step_frac": (balg * batch / (ms_dev / args.steps / 1000.0) / 1e9) / peak},
                "stage_ms": kern}
        if world == 1 and not args.skip_e2e and not args.no_dropin:
            # the user-facing path: Generator(args).run() over PNG files (decode + render + encode), SURVEY 8(f) rank 2
            try:
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import dropin_e2e
                import contextlib
                with contextlib.redirect_stdout(sys.stderr):          # Generator.run() prints progress like the reference: not on the JSON channel
                    line["dropin_png_e2e"] = dropin_e2e.measure(n_frames=args.dropin_frames, batch=batch)
            except Exception as e:                      # the headline numbers above stand on their own
                line["dropin_png_e2e"] = {"error": "%s: %s" % (type(e).__name__, e)}
        if world == 1 and not args.no_cpu_baseline:
            workers = max(1, min(host_cores(), 16))
            times, streaks = cpu_steps(1, 1, workers)
            fps = workers / times[0]
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": workers, "host_cores": host_cores(), "kind": "port",
                                    "sample": "%d processes x 1 frame (frames 0..%d of this batch, ~%d streaks/frame), oracle port, %.1f s wall" % (
                                        workers, workers - 1, int(np.mean(streaks)), times[0])}
        print(json.dumps(line))
    lanes.close()
    if world > 1:
        dist.destroy_process_group()


def run_gpu_c3(args):
    """BASELINE config C3: Cityscapes 2048x1024 frames rendered at 1024x512 (render_scale 2), 50 mm/h, ON-THE-FLY particle
    simulation.  A step = rr_simulate_records_device (simulator -> loader arithmetic -> in-frame filter -> RNG draws, records
    stay in HBM; only the 65 frame offsets come back) + rr_render_frames_device_io over 64 resident frames.  One GPU; prints
    one JSON line (kept under profiles/, not the driver's headline)."""
    import ctypes as C
    import torch
    from rain_rendering_b200 import _lib, api
    wl = synth.WORKLOADS["C3"]
    W, H, rs = wl["W"], wl["H"], 2
    cam = synth.CAMERAS[wl["dataset"]]
    batch = args.batch
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    db = synth.make_streak_db(0)
    frames = [synth.make_frame(W * rs, H * rs, i) for i in range(batch)]
    bgr = np.stack([f[0] for f in frames])
    d16 = np.stack([np.rint(synth.make_frame(W, H, i)[1] * 256.0).astype(np.uint16) for i in range(batch)])
    ctx = api.RainContext(0)
    ctx.set_streak_db(db.textures, db.ratios)
    ctx.set_camera(W, H, cam["cam_focal"], cam["cam_f_number"], cam["cam_exposure"], cam["cam_gain"], wl["fallrate"], 1.0, batch, render_scale=rs)
    stream_ptr = C.c_void_p()
    ctx.lib.rr_stream(ctx.h, C.byref(stream_ptr))
    stream = torch.cuda.ExternalStream(stream_ptr.value, device=dev)
    t_bgr, t_d = torch.from_numpy(bgr).to(dev), torch.from_numpy(d16.view(np.int16)).to(dev)
    t_out = torch.empty((batch, H, W, 3), dtype=torch.float32, device=dev)
    t_mask = torch.empty((batch, H, W), dtype=torch.float32, device=dev)
    t_u8 = torch.empty((batch, H, W, 3), dtype=torch.uint8, device=dev)
    t_idx = torch.empty((batch, H, W), dtype=torch.uint8, device=dev)
    pix = 4.65 * 1242.0 / (W * rs)                      # the synthetic optics scale the pixel pitch with the sensor width (synth.make_particles)
    n_streaks = []

    def step(k):
        ptr, offs, _ = ctx.simulate_records_device(k * batch, batch, W * rs, H * rs, wl["fallrate"], db.ratios, render_scale=rs, pix_size_um=pix,
                                                   exposure_ms=cam["cam_exposure"], seed=1)
        n_streaks.append(int(offs[-1]))
        ctx.render_frames_device(t_bgr.data_ptr(), t_d.data_ptr(), True, ptr, offs, t_out.data_ptr(), t_mask.data_ptr(), t_u8.data_ptr(), t_idx.data_ptr())

    for k in range(max(args.warmup, 3)):
        step(k)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    launches0 = ctx.kernel_launches()
    stage_ms = {k: 0.0 for k in ctx.timings()}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_streaks.clear()
    t0 = time.perf_counter()
    e0.record(stream)
    for k in range(args.steps):
        step(100 + k)
        for kk, v in ctx.timings().items():
            stage_ms[kk] += v
    e1.record(stream)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop()
    kern = {k: v / args.steps for k, v in stage_ms.items() if k not in ("h2d", "d2h", "total")}
    render_ms = stage_ms["total"] / args.steps
    line = {"metric": "rainy frames/sec at 2048x1024 -> 1024x512, 50mm/hr, on-the-fly simulation", "value": batch * args.steps / (ms / 1000.0), "unit": "frames/s",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C3 Cityscapes 2048x1024 (render 1024x512) 50mm/hr, on-the-fly particle simulation", "batch_frames_per_gpu": batch,
                       "streaks_per_frame": float(np.mean(n_streaks)) / batch,
                       "l2": "inputs larger than L2 (%.0f MB per step, no flush)" % ((bgr.nbytes + d16.nbytes) / 1e6)},
            "clocks": clocks, "gpu_launches": ctx.kernel_launches() - launches0,
            "sim_ms_per_step": ms / args.steps - render_ms, "render_ms_per_step": render_ms, "wall_ms_per_step": 1000.0 * wall / args.steps,
            "note": "simulation + record building on the device (4 launches, one 260-byte D2H of the frame offsets), then the render; no streak data crosses the host link",
            "stage_ms": kern}
    print(json.dumps(line))
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--lanes", type=int, default=2, help="independent contexts per GPU that take the steps in turn (api.RainLanes)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="C2", choices=["C2", "C3"], help="C2: the headline (default).  C3: Cityscapes with on-the-fly simulation, one GPU")
    ap.add_argument("--no-dropin", action="store_true", help="skip the drop-in Generator.run() PNG pass (dropin_png_e2e)")
    ap.add_argument("--dropin-frames", type=int, default=1024)
    ap.add_argument("--skip-e2e", action="store_true", help="profiling aid: only the device-resident arm (the JSON line is then incomplete)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "C3":
        run_gpu_c3(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
