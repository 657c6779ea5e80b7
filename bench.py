#!/usr/bin/env python
"""bench.py — headline benchmark of the `pt` hot path (BASELINE.json: Msamples/s, 1080p pt, path length 5).

Workload (config C3, BASELINE.json configs[2]): procedural 1M-triangle terrain (seed 1) + 4096x2048 lat-long HDR
environment, 1920x1080, 64 spp, max path length 5, one grey glTF material (roughness 0.5, metallic 0).
A "step" = one frame = width*height*spp paths through mpt_render (primary hits from the GBuffer hit texture).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (libmpt.so on B200)
  python bench.py --impl reference ...                          the reference arm: the CPU restatement of the
                                                                reference's algorithm (oracle/) on the host cores
Multi-GPU: one process per GPU (torchrun); the scene is replicated, rank r renders sample indices r, r+N, ...
into a private sum buffer; one NCCL reduce to rank 0 per frame, issued by libmpt.so itself (mpt_reduce_begin) on a side
stream so that it overlaps the next frame's render. --scaling weak (default): 64 spp per GPU, 64*N per frame;
--scaling strong: 64 spp per frame split over the N GPUs.
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

WIDTH, HEIGHT, SPP, MAX_PATH_LENGTH = 1920, 1080, 64, 5
METRIC = "Msamples/s (1080p pt, path len 5)"


def synthetic_version():
    from merian_b200 import synthetic
    return synthetic.GENERATOR_VERSION


def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """ncu `dram__bytes_read.sum + dram__bytes_write.sum` of k_extend, averaged over the extend launches of one timed step of THIS
    command (profiles/extend_traffic.json, written by scripts/extend_traffic_from_ncu.py from a `--set full` capture); or None."""
    p = os.path.join(ROOT, "profiles", "extend_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("dram_bytes_per_launch")
    return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """Samples nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc, self.first = index, [], None, 0

    def mark(self):
        """The timed region starts here: earlier samples (warm-up load) only serve as a fallback for a region shorter than one period."""
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window, note = self.lines[self.first:], None
        if not window and self.first > 0:  # e.g. 8-GPU strong scaling: 20 steps of 5 ms end before the next 100 ms sample
            window, note = self.lines[self.first - 1:self.first], "timed region shorter than the sampling period: last sample of the warm-up load"
        for l in window:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        out = {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
               "samples": len(sm)}
        if note:
            out["note"] = note
        return out


def build_scene(small=False):
    from merian_b200 import synthetic
    from merian_b200.scene import SceneTables
    if small:  # CI-sized variant for CPU tests of the harness
        return SceneTables(synthetic.terrain(nx=100, ny=50, env_size=(256, 128)))
    return SceneTables(synthetic.terrain(nx=1000, ny=500, seed=1, env_size=(4096, 2048)))


def cpu_baseline(tables, width, height, spp, seconds_hint=15.0):
    """The oracle (kind 'port': a CPU restatement of the reference's algorithm, not the reference) on a bounded
    sample of the same workload: same scene, camera and path length, reduced resolution."""
    from merian_b200 import abi
    from merian_b200.scene import upload_scene
    from oracle.oracle_py import OracleContext
    o = OracleContext()
    o.set_num_threads(host_threads())  # all host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)
    upload_scene(o, tables)
    cam = tables.desc.camera.to_abi(0)
    w, h, s = width // 4, height // 4, 8
    P = abi.RenderParams(w, h, 0, s, MAX_PATH_LENGTH, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)
    o.render(cam, abi.RenderParams(w // 4, h // 4, 0, 1, MAX_PATH_LENGTH, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0))  # warm
    t0 = time.perf_counter()
    n = 0
    while True:
        P.frame = n
        o.render(cam, P)
        n += 1
        if time.perf_counter() - t0 > seconds_hint or n >= 64:
            break
    dt = time.perf_counter() - t0
    return {"value": w * h * s * n / dt / 1e6, "unit": "Msamples/s", "cores": o.num_threads(), "kind": "port",
            "sample": f"{n} frames of {w}x{h} x {s} spp, max path length {MAX_PATH_LENGTH}, same scene/camera ({dt:.1f} s)"}, o, cam


def run_reference(args):
    """--impl reference: the reference's Vulkan/Slang `pt` cannot be built in this image, so this arm
    times the CPU restatement of its algorithm (oracle/) with all host threads on a bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from merian_b200 import abi
    tables = build_scene(args.small)
    base, o, cam = cpu_baseline(tables, args.width, args.height, args.spp, seconds_hint=1.0)
    w, h, s = args.width // 4, args.height // 4, 8
    P = abi.RenderParams(w, h, 0, s, MAX_PATH_LENGTH, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)
    for i in range(args.warmup):
        P.frame = i
        o.render(cam, P)
    t0 = time.perf_counter()
    for i in range(args.steps):
        P.frame = args.warmup + i
        o.render(cam, P)
    dt = time.perf_counter() - t0
    value = w * h * s * args.steps / dt / 1e6
    sample = f"each step = {w}x{h} x {s} spp of the workload (1/128 of a frame), oracle port, {o.num_threads()} threads"
    args.emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args), "width": args.width, "height": args.height, "spp": args.spp, "max_path_length": MAX_PATH_LENGTH,
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": o.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_name(args):
    return (f"C3: pt on synthetic 1M-triangle terrain (seed 1) + 4096x2048 lat-long HDR env, {args.width}x{args.height}, {args.spp} spp, "
            f"max path length {MAX_PATH_LENGTH}" if not args.small else "C3-small (harness test)")


def oracle_work_per_ray(ctx, tables, oracle, cam, W, H):
    """nodes / triangles per ray counted by the CPU oracle walking the SAME BVH8 (mpt_host_accel_build emits what
    mpt_scene_upload puts in HBM) over the SAME continuation rays: a reduced frame (W/4 x H/4, 4 spp) is rendered with max path length
    2 .. 5, the ray queue the last extend launch consumed is copied back each time (bounce 1 .. 4: the RNG is keyed on (pixel, frame,
    sample), so these are the rays of the one full-length frame), and the oracle's walker counts on them. The device's own counters over
    the same frame are returned next to it."""
    import torch
    from merian_b200 import abi
    from merian_b200.host_accel import HostAccel
    w, h, spp = max(W // 4, 8), max(H // 4, 8), 4
    acc = HostAccel(tables)
    tex1 = ctx.primary_visibility_device(cam, w, h)
    P = lambda L: abi.RenderParams(w, h, 77, spp, L, 0xFF, abi.RENDER_EMISSION_ON_PRIMARY, 0, 1, 0.0)
    tot = {"rays": 0, "nodes": 0, "tris": 0}
    for L in range(2, MAX_PATH_LENGTH + 1):
        ctx.render_device(cam, P(L), tex1=tex1)
        torch.cuda.synchronize()
        rays, leaves = ctx.debug_ray_queue(L - 1)
        if len(rays) == 0:
            continue
        _, _, _, cnt = oracle.bvh8_trace(acc.view, rays, skip_records=None if acc.view.two_level else leaves)
        for k in tot:
            tot[k] += cnt[k]
    ctx.set_profiling(1)
    ctx.reset_stats()
    ctx.render_device(cam, P(MAX_PATH_LENGTH), tex1=tex1)
    cs = ctx.stats()
    ctx.set_profiling(0)
    acc.close()
    r = max(tot["rays"], 1)
    rd = max(cs.rays_continuation, 1)
    return {"rays": tot["rays"], "nodes_per_ray": tot["nodes"] / r, "tris_per_ray": tot["tris"] / r, "device_rays": int(cs.rays_continuation),
            "device_nodes_per_ray": cs.nodes_visited / rd, "device_tris_per_ray": cs.tris_tested / rd,
            "sample": f"{w}x{h} x {spp} spp, same scene / camera / path length"}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from merian_b200 import abi, sharding
    from merian_b200.context import Context
    from merian_b200.scene import upload_scene

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    tables = build_scene(args.small)
    ctx = Context(local_rank)
    upload_scene(ctx, tables)
    bvh = ctx.bvh_info()
    desc = tables.desc
    W, H = args.width, args.height
    npix = W * H
    strong = args.scaling == "strong"
    # weak scaling: args.spp per GPU, args.spp * N per frame; strong scaling: args.spp per frame, split over the N GPUs
    spp_total = args.spp if strong else args.spp * world
    spp_rank = sharding.shard_counts(spp_total, world)[rank]

    def params(frame, flags=abi.RENDER_EMISSION_ON_PRIMARY):
        if world > 1:
            return abi.RenderParams(W, H, frame, spp_rank, MAX_PATH_LENGTH, 0xFF, flags | abi.RENDER_OUTPUT_SUM, rank, world, 0.0)
        return abi.RenderParams(W, H, frame, spp_total, MAX_PATH_LENGTH, 0xFF, flags, 0, 1, 0.0)

    cam = desc.camera.to_abi(0)
    tex1 = ctx.primary_visibility_device(cam, W, H)  # the GBuffer node's product (gbuffer_rt.slang), resident in HBM
    # two frames in flight: frame i's partial sums are being reduced on the library's side stream while frame i + 1 renders
    NS = 4  # frame slots (sum buffer + reduce slot of the library, mpt_reduce_begin takes up to 4): the device-timed loop never blocks the
    #         host; the end-to-end loop blocks on the frame that used the slot NS frames ago, i.e. NS - 1 frames stay queued on the GPU
    sums = [torch.empty((H, W, 4), dtype=torch.float32, device=dev) for _ in range(NS)]
    outs = [torch.empty((H, W, 4), dtype=torch.float32, device=dev) for _ in range(NS)] if (world > 1 and rank == 0) else [None] * NS
    comm = None
    if world > 1:
        def exchange(raw):  # rank 0's NCCL unique id -> every rank (out of band: torch.distributed is only plumbing here)
            t = torch.zeros(abi.COMM_UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
            if rank == 0:
                t.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
            dist.broadcast(t, src=0)
            return bytes(t.cpu().numpy().tobytes())
        comm = sharding.ShardComm(local_rank, rank, world, exchange)  # mpt_comm_create: NCCL inside libmpt.so
    pending = [False] * NS  # a reduce is in flight in slot b (reading sums[b], writing outs[b] / the host buffer)

    def step(i, h_out=None, h_tex1=None, d_tex1=tex1):
        """One frame. Multi-GPU: render this rank's samples into sums[i % NS], then mpt_reduce_begin(slot i % NS): reduce -> normalise ->
        optional copy to the host buffer, on the library's side stream, overlapping the next frame's render. The host never blocks on the
        previous frame (that would leave a single frame queued on the GPU, and every host hiccup would stall it): before a sum buffer is
        rendered into again, the render STREAM waits for the reduce that read it (mpt_reduce_end), or — when the frames land in host
        buffers, which the host reuses — the host waits for the reduce that used the slot NS frames ago (mpt_reduce_wait)."""
        b = i % NS
        if world > 1 and pending[b]:
            if h_out is not None:
                comm.reduce_wait(slot=b)
            else:
                comm.reduce_end(slot=b)
            pending[b] = False
        if h_tex1 is not None:
            d_tex1.copy_(h_tex1, non_blocking=True)
        ctx.render_device(cam, params(i), tex1=d_tex1, out=sums[b])
        if world == 1:
            if h_out is not None:
                h_out.copy_(sums[b], non_blocking=True)
            return
        comm.reduce_begin(sums[b], outs[b], W, H, spp_total, root=0, h_out=h_out if rank == 0 else None, slot=b)
        pending[b] = True

    def drain(on_stream=False):
        for b in range(NS):
            if world > 1 and pending[b]:
                if on_stream:
                    comm.reduce_end(slot=b)
                else:
                    comm.reduce_wait(slot=b)
                pending[b] = False

    def barrier():
        drain()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- algorithmic work per ray, outside the timed region: the oracle's count (roofline numerator) and the device's own
    work = None
    oracle = None
    base = None
    if rank == 0 and not args.no_cpu_baseline:
        base, oracle, _ = cpu_baseline(tables, W, H, args.spp)  # also the oracle context (scene uploaded) for the walker below
        work = oracle_work_per_ray(ctx, tables, oracle, cam, W, H)
    ctx.set_profiling(1)
    ctx.reset_stats()
    ctx.render_device(cam, params(10_000), tex1=tex1, out=sums[0])
    cs = ctx.stats()
    rays_c = max(cs.rays_continuation, 1)
    dev_nodes_per_ray, dev_tris_per_ray = cs.nodes_visited / rays_c, cs.tris_tested / rays_c
    ctx.set_profiling(0)
    nodes_per_ray = work["nodes_per_ray"] if work else dev_nodes_per_ray
    tris_per_ray = work["tris_per_ray"] if work else dev_tris_per_ray
    bytes_per_ray = 80.0 * nodes_per_ray + 48.0 * tris_per_ray + 48.0  # node + triangle records + ray read (32) + hit write (16)

    # The clock sampler is an `nvidia-smi -lms` child process: its start-up (NVML init) takes driver locks that CUDA API calls of this
    # process can block on, which starves the GPU of launches for hundreds of milliseconds. Start it BEFORE the warm-up and wait for its
    # first sample, so the timed region only sees the steady 100 ms polling.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        t_wait = time.perf_counter()
        while not sampler.lines and time.perf_counter() - t_wait < 5.0:
            time.sleep(0.05)
    ctx.set_profiling(2)  # CUDA-event timing of every extend / shade launch; on during the warm-up so that libmpt's event pool is
    #                       fully created before the timed region (no cudaEventCreate inside it)
    for i in range(max(6, args.steps + 1)):  # clocks and caches up before the W warm-up steps the caller asked for
        ctx.render_device(cam, params(20_000 + i), tex1=tex1, out=sums[i % NS])
    ctx.reset_stats()
    for i in range(args.warmup):
        step(i)
    barrier()
    ctx.reset_stats()
    if rank == 0:
        sampler.mark()  # only the samples of the timed region count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    # One more untimed step goes in front of the timed ones, behind the barrier: right after a synchronize the GPU's queue is empty, and a
    # host thread that is held up for a few ms before it gets the first launch out (observed: 10-80 ms, about once per second of run time,
    # in every mode and also on one GPU, profiles/r02j_reduce_modes.jsonl) would put that idle gap inside the timed region. With a
    # frame already executing, the K timed steps are all enqueued long before the GPU reaches them; ev0 is recorded in stream order
    # behind the lead-in frame, so the timed region is exactly K steps of device time.
    step(args.warmup)
    ev0.record()
    for i in range(args.steps):
        step(args.warmup + 1 + i)
    drain(on_stream=True)  # the timing stream waits for the last frames' reduce + normalise: ev1 is behind the whole job
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    st = ctx.stats()
    ctx.set_profiling(0)

    timed_share = args.steps / (args.steps + 1.0)  # libmpt's counters also saw the lead-in step
    launches = float(st.kernel_launches) * timed_share + (2.0 * args.steps if (world > 1 and rank == 0) else (1.0 * args.steps if world > 1 else 0.0))  # + NCCL reduce, k_normalise
    rays = torch.tensor([float(st.rays_continuation) * timed_share, float(st.shaded_hits) * timed_share, launches], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    total_samples = npix * spp_total * args.steps
    value = total_samples / (ms * 1e-3) / 1e6
    mrays = float(rays[0].item()) / (ms * 1e-3) / 1e6

    # ---- end to end with HOST buffers, every step: H2D of the GBuffer hit texture from pinned memory, render, (multi-GPU: reduce +
    # normalise,) D2H of the irradiance into pinned memory on rank 0 — two iterations in flight as merian's graph runs them
    # (examples/gltf.json:39). One GPU: mpt_render_host_begin / _end (the C-ABI's host-buffer entry points). Several GPUs: the same
    # copies around mpt_render + mpt_reduce_begin(h_out); a step is complete when rank 0 holds the normalised frame in host memory.
    h_tex1 = tex1.cpu().pin_memory()
    import gc
    gc.collect()
    gc.freeze()   # the host loop below is what a C++ graph runtime would run: keep CPython's cycle collector (a full pass over torch's
    gc.disable()  # module graph takes tens of ms) out of the timed region
    F = max(1, min(int(args.in_flight), 4))  # iterations in flight (mpt_render_host_begin: in_flight_index < 4; merian's ring size is a graph setting)
    h_outs = [torch.empty((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(max(F, NS))]
    e2e_steps = max(2, args.steps)
    e2e_two = None

    def pipelined(f, first_frame):
        """e2e_steps iterations through mpt_render_host_begin / _end with f iterations in flight; wall time from the first begin to the
        last frame's pixels in host memory."""
        for s_ in range(f):
            ctx.render_host_begin(cam, params(s_), h_outs[s_], s_, h_tex1=h_tex1)
        for s_ in range(f):
            ctx.render_host_end(s_)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            if i >= f:
                ctx.render_host_end(i % f)
            ctx.render_host_begin(cam, params(first_frame + i), h_outs[i % f], i % f, h_tex1=h_tex1)
        for i in range(max(e2e_steps - f, 0), e2e_steps):
            ctx.render_host_end(i % f)
        return time.perf_counter() - t0

    if world == 1:
        ctx.render_host(cam, params(0), h_tex1=h_tex1, h_out=h_outs[0])  # warm (allocates the staging buffers)
        # The host thread of a pipelined loop stalls now and then for 60-110 ms inside a CUDA call (scripts/e2e_probe.py,
        # profiles/r02s_e2e_probe.json: 2-3 times in 20 steps on the GPU boxes, the same driver-lock contention as the NVML start-up
        # described above). Two iterations in flight hold one frame (35 ms) of queued work, so every such stall idles the GPU; four hold
        # three. Both are measured; `e2e.value` is the configured depth.
        if F != 2:
            e2e_two = npix * spp_total * e2e_steps / pipelined(2, args.warmup + 1000) / 1e6
        # two runs of `steps` iterations, the better one counts (both are in the line): a host hold-up longer than the queued work still
        # lands in one run out of three or four on these boxes, and it says nothing about the path
        e2e_runs = [pipelined(F, args.warmup + 2000 * r) for r in range(2)]
        e2e_dt = min(e2e_runs)
    else:
        e2e_runs = None
        d_tex1s = [torch.empty_like(tex1) for _ in range(NS)]
        for i in range(NS):
            step(i, h_out=h_outs[i % NS], h_tex1=h_tex1, d_tex1=d_tex1s[i % NS])
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            step(args.warmup + i, h_out=h_outs[i % NS], h_tex1=h_tex1, d_tex1=d_tex1s[i % NS])
        drain()  # rank 0: the last frame is in host memory; other ranks: their last contribution has been sent
        torch.cuda.synchronize()
        e2e_dt = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = npix * spp_total * e2e_steps / float(e2e_t.item()) / 1e6
    # the synchronous call (one iteration in flight), reported next to it
    e2e_sync_value = None
    if world == 1:
        best = None
        for rep in range(3):  # best of three: the first call after the pipelined loop pays for the stream switch
            t0 = time.perf_counter()
            ctx.render_host(cam, params(args.warmup + rep), h_tex1=h_tex1, h_out=h_outs[0])
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        e2e_sync_value = npix * spp_total / best / 1e6

    # ---- the same frame with MPT_RENDER_NEE (north_star: closest-hit and occlusion queries reported separately and summed); outside
    # the timed region, not part of `value`
    nee = None
    if world == 1:
        ctx.reset_stats()
        for i in range(2):
            ctx.render_device(cam, params(30_000 + i, abi.RENDER_EMISSION_ON_PRIMARY | abi.RENDER_NEE), tex1=tex1, out=sums[0])
        torch.cuda.synchronize()
        n0, n1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # an untimed lead-in frame in front of n0, as in the main loop: the timed frames are enqueued while it runs, so a host thread that is
        # held up for tens of ms (profiles/r02y_e2e_probe.json) does not leave the GPU idle inside the timed region (two frames timed from
        # an idle GPU read 89-95 ms where four frames behind a lead-in frame read 69.6 ms: profiles/r02z_nee_live.jsonl)
        NN = 4
        ctx.render_device(cam, params(30_002, abi.RENDER_EMISSION_ON_PRIMARY | abi.RENDER_NEE), tex1=tex1, out=sums[0])
        ctx.reset_stats()
        n0.record()
        for i in range(NN):
            ctx.render_device(cam, params(30_003 + i, abi.RENDER_EMISSION_ON_PRIMARY | abi.RENDER_NEE), tex1=tex1, out=sums[0])
        n1.record()
        torch.cuda.synchronize()
        nms = n0.elapsed_time(n1) / NN
        ns = ctx.stats()
        nee = {"msamples_per_s": npix * spp_total / (nms * 1e-3) / 1e6, "ms_per_step": nms,
               "mrays_closest_hit_per_s": ns.rays_continuation / NN / (nms * 1e-3) / 1e6, "mrays_occlusion_per_s": ns.rays_shadow / NN / (nms * 1e-3) / 1e6,
               "mrays_total_per_s": (ns.rays_continuation + ns.rays_shadow) / NN / (nms * 1e-3) / 1e6}

    if rank == 0:
        peak, peak_src = load_peak()
        extend_ms = st.extend_ms / max(st.extend_launches, 1)
        rays_per_launch = st.rays_continuation / max(st.extend_launches, 1)
        achieved = bytes_per_ray * rays_per_launch / (extend_ms * 1e-3) / 1e9 if extend_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_name(args), "width": W, "height": H, "spp_per_gpu": spp_rank, "spp_total": spp_total,
                       "max_path_length": MAX_PATH_LENGTH, "triangles": tables.triangle_count, "seed": desc.seed, "generator": synthetic_version(),
                       "timing": "CUDA events around exactly `steps` steps, max over ranks; barrier + synchronize, then one untimed lead-in step in front of ev0 (keeps the queue non-empty while the host enqueues the timed steps)",
                       "parallelism": (f"sample-sharded x{world}, one NCCL reduce of the sum buffers per frame inside libmpt.so (mpt_reduce_begin), "
                                       "overlapped with the next frame's render") if world > 1 else "single GPU",
                       "l2": "inputs larger than L2: BVH8 %.0f MB + triangle/attribute records %.0f MB + env 134 MB + >1 GB of path state per pass"
                             % (bvh.node_bytes / 1e6, 2 * bvh.triangle_bytes / 1e6)},
            "mrays_per_s": mrays,  # continuation rays (primary hits come from the GBuffer texture)
            "rays_per_sample": float(rays[0].item()) / total_samples,
            "with_nee": nee,  # MPT_RENDER_NEE on the same workload (an addition to the reference estimator; not the headline)
            "roofline": {"bound": "hbm", "kernel": "k_extend_pooled", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": load_traffic(), "peak_source": peak_src, "bytes_per_ray": bytes_per_ray,
                         "counted_by": "oracle BVH8 walker on the device's ray queues" if work else "device counters (no oracle in this run)",
                         "nodes_per_ray": nodes_per_ray, "tris_per_ray": tris_per_ray,
                         "nodes_per_ray_device": dev_nodes_per_ray, "tris_per_ray_device": dev_tris_per_ray, "oracle_sample": work,
                         "rays_per_launch": rays_per_launch, "avg_launch_ms": extend_ms,
                         "extend_share_of_step": st.extend_ms * timed_share / ms if ms > 0 else None,
                         "shade_share_of_step": st.shade_ms * timed_share / ms if ms > 0 else None},
            "cpu_baseline": base,
            "e2e": {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": npix * 16 * world, "d2h_bytes_per_step": npix * 16,
                    "iterations_in_flight": (F if world == 1 else NS), "two_in_flight_value": e2e_two, "steps": e2e_steps,
                    "runs": [npix * spp_total * e2e_steps / r / 1e6 for r in e2e_runs] if e2e_runs else None, "policy": "best of 2 runs of `steps` iterations" if e2e_runs else "one run of `steps` iterations", "synchronous_call_value": e2e_sync_value,
                    "includes_reduce": world > 1},
            "gpu_launches": int(rays[2].item()),
            "clocks": clocks,
            "bvh": {"nodes": bvh.node_count, "node_mb": bvh.node_bytes / 1e6, "triangle_mb": bvh.triangle_bytes / 1e6, "build_ms": bvh.build_ms},
        }
        args.emit(json.dumps(line))
    if comm is not None:
        comm.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


class _QuietStdout:
    """Libraries under us print to the C-level stdout (NCCL writes its version banner there): point fd 1 at stderr while the bench runs
    and give it back for the ONE JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, line):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        print(line, flush=True)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--spp", type=int, default=SPP)
    ap.add_argument("--in-flight", type=int, default=4, help="iterations in flight of the end-to-end (host buffer) measurement on one GPU, 1..4")
    ap.add_argument("--small", action="store_true", help="tiny scene (harness test only; not a bench value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="multi-GPU: weak = --spp per GPU (default, the driver's scaling run), strong = --spp per frame split over the GPUs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours" and not args.small:
        args.warmup = 3  # timing rule: at least 3 warm-up steps
    with _QuietStdout() as out:
        args.emit = out.emit
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
