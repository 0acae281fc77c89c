#!/usr/bin/env python
"""bench.py — the reference's headline metric (BASELINE.json): frames/s of the rasterization hot
path at 3840x2160 on a synthetic scene of random translucent circles, with the fraction of the
measured HBM roofline.

  python bench.py --gpus 1 --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference --gpus 1 ...            the reference's own kernels (Kernels.cl compiled
                                                           for the host, oracle/_ref, OpenMP over the
                                                           NDRange) on the host cores; the restated port
                                                           in oracle/ only if oracle/_ref is not built
  torchrun ... bench.py --gpus N ...                       N > 1: one 16384^2 canvas partitioned
                                                           into tile-row strips, one rank per GPU,
                                                           strips gathered on rank 0 (strong scaling)

A "step" is one frame: tile binning + threshold generation + sort + sweep/compositing, producing
the BGRA8 bitmap.  `value` times it with the scene resident in HBM and the frame left in HBM;
`e2e` times the public call a client makes (host buffers in, host bitmap out).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (scene factory name, description)
    "S4": ("s4", "S4: 100,000 random translucent circles (1.7M quadratic curves), 3840x2160, seed 0x5EED0004"),
    "S4b": ("s4b", "S4b: 6,250 random translucent circles (100k curves), 3840x2160, seed 0x5EED004B"),
    "S5": ("s5", "S5: 62,500 random translucent circles r 20-200 (1M curves), 16384x16384, seed 0x5EED0005"),
    "S5b": ("s5b", "S5b: 1,000,000 random translucent circles r 5-10, 16384x16384, seed 0x5EED005B"),
}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                parts = [p.strip() for p in out.split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for n, v in zip(names, parts[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=3)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def make_scene(workload):
    from gudni_b200 import scenes
    return getattr(scenes, WORKLOADS[workload][0])()


def algorithmic_bytes(scene, n_tiles, n_shape_refs, rows=None):
    """A(frame), SURVEY.md §8(d)."""
    rows = scene.height if rows is None else rows
    return (scene.geometry.nbytes + 16 * n_shape_refs + 32 * n_tiles + 16 * len(scene.substances) +
            24 * len(scene.picture_uses) + scene.picture_bytes.nbytes + 4 * scene.width * rows)


# ---------------------------------------------------------------------------------------------------
# reference arm: the reference's own algorithm (three phases per job over the CPU tile tree) on the
# host cores.  The Haskell host layer cannot be built here (no GHC, no OpenCL runtime), but the kernel
# file can: oracle/_ref/libgudni_ref.so is Kernels.cl compiled by g++ against a small OpenCL-C
# compatibility header (oracle/refbuild/), its NDRange played by an OpenMP loop with the reference's
# scratch layout and three passes per job (kind "reference").  Without that library (no reference tree
# at build time) the restated port in oracle/ stands in (kind "port").  The tile tree is the restated
# one in both cases (Raster/TileTree.hs is Haskell).
# ---------------------------------------------------------------------------------------------------
def cpu_kind():
    from oracle import oracle
    return "reference" if oracle.reference_lib() is not None else "port"


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arms are meant to use every host core, so the
    OpenMP pools of the oracle and of the compiled reference kernels are sized explicitly."""
    from oracle import oracle
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0)) or n
    except (AttributeError, OSError):
        pass
    oracle.lib().gudni_oracle_set_threads(n)
    ref = oracle.reference_lib()
    if ref is not None:
        ref.gudni_ref_set_threads(n)
    return n


def oracle_frame_sampler(scene, budget_s=12.0, kind=None):
    """Returns f() -> (seconds per FULL frame, sample description).  Frames that would take longer
    than `budget_s` are sampled: the tile tree is built for the whole scene, then an evenly spread
    subset of the jobs is rasterized and the raster time is scaled by jobs_total / jobs_sampled."""
    from oracle import oracle
    use_ref = (kind or cpu_kind()) == "reference"

    t0 = time.perf_counter()
    jobs = oracle.build_raster_jobs(scene)
    t_tree = time.perf_counter() - t0
    probe = jobs[len(jobs) // 2: len(jobs) // 2 + 1]
    t0 = time.perf_counter()
    oracle.raster_jobs(scene, probe, taps=False, reference=use_ref)
    t_probe = time.perf_counter() - t0
    est = t_tree + t_probe * len(jobs)
    stride = max(1, int(np.ceil(est / budget_s)))
    picked = jobs[stride // 2::stride] if stride > 1 else jobs
    desc = (f"tile tree for the whole scene + {len(picked)} of {len(jobs)} raster jobs "
            f"(one job in {stride} of {scene.name}), raster time scaled by {len(jobs)}/{len(picked)}"
            if stride > 1 else f"one full frame of {scene.name} (tile tree + {len(jobs)} raster jobs)")

    def run():
        t0 = time.perf_counter()
        js = oracle.build_raster_jobs(scene)
        t1 = time.perf_counter()
        sel = js[stride // 2::stride] if stride > 1 else js
        oracle.raster_jobs(scene, sel, taps=False, reference=use_ref)
        t2 = time.perf_counter()
        return (t1 - t0) + (t2 - t1) * len(js) / len(sel)

    run.sampled = stride > 1
    run.est_seconds = est / stride if stride > 1 else est     # wall clock of one sampled step
    return run, desc


def level3_reading(r, scene, dscene, stream, flush, torch):
    """Secondary reading, not part of `value`: the same frame entered one level higher,
    from the scene BEFORE serialisation — outlines + transformer stacks (S4: one 17-pair outline, 100,000
    placements) — with the strands built on the GPU.  Device-resident (CUDA events, L2 flushed) and end to end
    with host buffers; beside it the harness's single-core serialisation of the same scene, which is what the
    Haskell front end does per frame today."""
    dscene.put_outlines()
    ms, strands_ms = [], []
    for i in range(3 + 10):
        flush.fill_(i & 0xFF)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        r.frame_begin_device_outlines(dscene, i)
        r.raster_outlines_device(dscene)
        _, st = r.frame_end(want_image=False)
        e1.record(stream)
        torch.cuda.synchronize()
        if i >= 3:
            ms.append(e0.elapsed_time(e1))
            strands_ms.append(st.ms_strands)
    host_img = np.empty((scene.height, scene.width), dtype=np.uint32)
    arrays = [np.ascontiguousarray(a) for a in scene.raw] + [scene.substances, host_img]
    scene.raw = tuple(arrays[:4])
    for a in arrays:
        r.host_register(a)
    e2e = []
    for i in range(2 + 5):
        t0 = time.perf_counter()
        r.raster_outlines(i, scene, out=host_img)
        if i >= 2:
            e2e.append(time.perf_counter() - t0)
    for a in arrays:
        r.host_unregister(a)
    from gudni_b200 import scenes as scene_factories
    t0 = time.perf_counter()
    scene_factories.s4()
    t_host = time.perf_counter() - t0
    h2d = sum(a.nbytes for a in arrays[:5]) + scene.picture_bytes.nbytes + scene.picture_uses.nbytes
    return {"value": 1e3 / float(np.mean(ms)), "unit": "frames/s", "ms_per_step": float(np.mean(ms)),
            "ms_strands": float(np.mean(strands_ms)),
            "strand_kernels": "strand_measure_kernel + strand_scan_kernel + strand_emit_kernel",
            "strand_bytes": {"in": int(sum(a.nbytes for a in arrays[:4])), "out": int(scene.geometry.nbytes + scene.entries.nbytes)},
            "strand_gbs": (sum(a.nbytes for a in arrays[:4]) + scene.geometry.nbytes + scene.entries.nbytes) / (float(np.mean(strands_ms)) * 1e-3) / 1e9,
            "e2e": {"value": 1.0 / float(np.mean(e2e)), "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(host_img.nbytes)},
            "host_serialisation_ms_1_core": t_host * 1e3}


def opencl_reference_on_this_gpu(workload, timeout_s=90):
    """Side measurement, not an arm: the reference's Kernels.cl, verbatim, under the OpenCL runtime of the
    GPU box (NVIDIA OpenCL on the same B200), same scene, same raster jobs — oracle/refbuild/ocl_run.py in a
    child process with a time limit.  Needs oracle/_ref/ocl_program.bin (built where /root/reference
    exists) and libnvidia-opencl on the box; says why when it cannot run."""
    import tempfile
    blob = os.path.join(ROOT, "oracle", "_ref", "ocl_program.bin")
    if not os.path.exists(blob):
        return {"unavailable": "oracle/_ref/ocl_program.bin not built (no reference tree at build time)"}
    scene_key = WORKLOADS[workload][0]
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "ocl.json")
        cmd = [sys.executable, os.path.join(ROOT, "oracle", "refbuild", "ocl_run.py"), "--out", out,
               "--scenes", scene_key, "--variants", "reference", "--frames", "3"]
        try:
            subprocess.run(cmd, timeout=timeout_s, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False)
            with open(out) as f:
                log = json.load(f)
            if scene_key not in log.get("results", {}).get("reference", {}):
                last = log["steps"][-1] if log.get("steps") else {}
                return {"unavailable": (str(last.get("error") or last.get("msg") or "no result"))[:200]}
            res = log["results"]["reference"][scene_key]
        except Exception as e:  # noqa: BLE001 - a side measurement never fails the bench
            return {"unavailable": repr(e)[:200]}
    return {"value": 1.0 / res["kernel_seconds"], "unit": "frames/s", "kernel_ms": res["kernel_seconds"] * 1e3,
            "frame_ms_with_scratch_alloc": res["frame_seconds"] * 1e3, "frames_timed": res["frames_timed"],
            "device": log.get("device", {}).get("name"), "runtime": log.get("device", {}).get("version"),
            "options": "-cl-fast-relaxed-math -cl-strict-aliasing (OpenCL/Setup.hs:126-129)",
            "timing": "wall clock around the three launches of every job, clFinish on both sides, best frame",
            "pixels_vs_oracle": {"exact_rate": res["exact_rate"], "max_channel_diff": res["max_channel_diff"],
                                 "threshold_counts_equal": res["threshold_counts_equal"]}}


CPU_NOTES = {
    "reference": "the reference's own Kernels.cl compiled for the host (g++ -O3, IEEE f32, OpenMP over the NDRange, "
                 "scratch layout and three passes per job as in OpenCL/CallKernels.hs:88-179); tile tree restated "
                 "(Haskell); not PoCL",
    "port": "restated-reference CPU (oracle/, OpenMP): oracle/_ref was not built (no reference tree at build time)",
}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle
    cores = use_all_host_cores()
    scene = make_scene(args.workload)
    kind = cpu_kind()
    # The whole --steps/--warmup run has to end within a few minutes on the host cores: a step is one frame, or —
    # when a frame takes longer than the per-step budget — a bounded sample of its raster jobs scaled to the frame
    # ("sampled": true).  At most `max_steps` steps are timed however many were asked for; "steps" reports what ran.
    total_budget = 150.0
    budget = float(np.clip(total_budget / (args.steps + 1), 2.0, 12.0))
    run, desc = oracle_frame_sampler(scene, budget_s=budget, kind=kind)
    max_steps = max(1, int(total_budget / max(run.est_seconds, 1e-3)) - 1)
    steps = max(1, min(args.steps, max_steps))
    warmup = min(args.warmup, 1)
    for _ in range(warmup):
        run()
    t_wall = time.perf_counter()
    times = [run() for _ in range(steps)]
    t_wall = time.perf_counter() - t_wall
    t = float(np.mean(times))
    value = 1.0 / t
    line = {
        "impl": "reference", "metric": "frames/s", "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": steps, "steps_requested": args.steps, "warmup": warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "sampled": bool(run.sampled), "wall_seconds_timed": t_wall,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][1], "canvas": [scene.width, scene.height],
                   "spec": "G=256 MAXT=1024 maxStrandsPerTile=1022 MAXSHAPE=127"},
        "mpixel_per_s": scene.width * scene.height * value / 1e6,
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": desc,
                         "note": CPU_NOTES[kind]},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# native arm
# ---------------------------------------------------------------------------------------------------
def run_native(args, rank, world, local_rank):
    import torch
    from gudni_b200.raster import DeviceScene, setup_rasterizer
    from gudni_b200.strips import StripRenderer

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the CUDA path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    scene = make_scene(args.workload)
    r = setup_rasterizer(local_rank)
    stream = torch.cuda.current_stream()
    r.set_stream(stream.cuda_stream)
    strips = StripRenderer(r, scene, rank, world, dist, mode="p2p" if args.gather == "p2p" else "nccl")
    dscene = DeviceScene(r, scene, entries=strips.entries)
    pipelined = world > 1 and args.gather in ("nccl", "auto") and args.pipeline
    if pipelined:
        strips.prepare_chunks(dscene.put_entries, chunk_rows=args.chunk_rows)
    render = (lambda f: strips.render_pipelined(f, dscene)) if pipelined else (lambda f: strips.render(f, dscene))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- value: inputs resident in HBM, frame left in HBM -----------------------------------------
    single = None
    solo_canvas = None
    if world > 1:
        # the same workload on ONE GPU, measured by rank 0 in this run, so the strong-scaling ratio
        # can be read off this line alone (bench.py --gpus 1 measures the 4K scene, not this one)
        if rank == 0:
            solo = StripRenderer(r, scene, 0, 1, None)
            dsolo = DeviceScene(r, scene)
            for i in range(2):
                solo.render(i, dsolo)
            t = []
            for i in range(3):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); solo.render(2 + i, dsolo); e1.record(stream)
                torch.cuda.synchronize()
                t.append(e0.elapsed_time(e1))
            single = {"ms_per_step": float(np.mean(t)), "value": 1e3 / float(np.mean(t)), "unit": "frames/s"}
            solo_canvas = solo.canvas.clone()      # the whole frame from ONE GPU: what the gathered canvas must equal
            solo.close(); dsolo.free(); del solo
        barrier()
    for i in range(args.warmup):
        render(i)
    gather_order = None
    if world > 1 and not pipelined and not args.no_rebalance:
        # Feedback partition (contiguous tile-row strips stay): a few rounds of "measure every rank's strip, cut
        # the canvas again" under each way of getting the strips to the presenting rank, keeping the fastest:
        #   nccl/early  the presenting rank posts its receives before its own strip; strips queue on its inbound
        #               links in the order their ranks finish, so each strip is charged the transfer of everything
        #               from its first row down (ranks finish staggered); the receive kernel holds some SMs meanwhile
        #   nccl/late   receives after the presenting rank's strip: nothing arrives before it is done, so it is charged
        #               the transfer of all the other strips too
        #   p2p         the presenting rank's canvas is mapped into every process (CUDA IPC) and the accumulate kernel
        #               stores its finished 128-byte rows there directly over NVLink: the transfer rides along with
        #               the rasterization, a barrier closes the frame; strips are balanced on kernel time alone
        from gudni_b200.strips import rebalance_rows
        from gudni_b200 import multi as multi_abi
        row_ms = 4.0 * scene.width * r.spec.max_tile_size / args.gather_gbs / 1e9 * 1e3
        start_rows = list(strips.rows)

        def frame_ms(n=3):
            t = []
            for i in range(n):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); render(i); e1.record(stream)
                barrier()
                t.append(e0.elapsed_time(e1))
            m = torch.tensor([float(np.mean(t))], dtype=torch.float64, device="cuda")
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
            return float(m.item())

        def rebuild(rows, mode, early):
            nonlocal strips, dscene, render
            strips.close(); dscene.free()
            strips = StripRenderer(r, scene, rank, world, dist, mode=mode, rows=rows, early_receives=early)
            dscene = DeviceScene(r, scene, entries=strips.entries)
            render = lambda f: strips.render(f, dscene)
            for i in range(2):
                render(i)

        if args.gather == "auto":
            candidates = [("nccl", "early"), ("nccl", "late"), ("p2p", None)]
        elif args.gather == "p2p":
            candidates = [("p2p", None)]
        else:
            candidates = [("nccl", o) for o in (["early", "late"] if args.gather_order == "auto" else [args.gather_order])]
        tried = {}
        for mode, order in candidates:
            rebuild(start_rows, mode, order == "early")
            for it in range(args.rebalance_rounds):
                st = getattr(strips, "last_stats", None)
                mine = torch.tensor([st.ms_raster + st.ms_bin if st is not None else 0.0], dtype=torch.float64, device="cuda")
                allr = [torch.zeros_like(mine) for _ in range(world)]
                dist.all_gather(allr, mine)
                times = [float(x.item()) for x in allr]
                if mode == "p2p":     # kernel time alone: the C ABI's own rebalancer (gudni_b200_rebalance_rows)
                    new_rows = multi_abi.rebalance_rows(strips.rows, times, scene.height, r.spec.max_tile_size)
                else:
                    new_rows = rebalance_rows(strips.rows, times, scene.height, r.spec.max_tile_size, 0, 0.0, row_ms,
                                              late_receives=(order == "late"))
                if new_rows == strips.rows:
                    break
                rebuild(new_rows, mode, order == "early")
            tried[(mode, order)] = (frame_ms(), list(strips.rows))
        best = min(tried, key=lambda k: tried[k][0])
        if len(tried) > 1 or strips.rows != tried[best][1]:
            rebuild(tried[best][1], best[0], best[1] == "early")
        gather_order = {"gather": best[0], "order": best[1],
                        "warmup_ms": {f"{k[0]}/{k[1]}" if k[1] else k[0]: round(v[0], 3) for k, v in tried.items()}}
    parity_ok = None
    if world > 1:
        # pixel evidence for the scaling record: the canvas gathered from N GPUs against the single-GPU frame
        canvas = render(args.warmup)
        barrier()
        if rank == 0:
            parity_ok = bool(torch.equal(canvas, solo_canvas))
            del solo_canvas
    launches0 = r.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms, raster_ms, bin_ms = [], [], []
    for i in range(args.steps):
        flush.fill_(i & 0xFF)                      # L2 flush between timed iterations (not timed)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        render(args.warmup + i)
        e1.record(stream)
        barrier()
        step_ms.append(e0.elapsed_time(e1))
        st = getattr(strips, "last_stats", None)
        if st is not None:
            raster_ms.append(st.ms_raster)
            bin_ms.append(st.ms_bin)
    clocks = sampler.stop()
    launches = r.launch_count() - launches0
    per_rank = None
    if dist is not None:
        mine = torch.tensor([float(np.mean(raster_ms)) if raster_ms else 0.0, float(np.mean(bin_ms)) if bin_ms else 0.0],
                            dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[round(float(x[0]), 3), round(float(x[1]), 3)] for x in allr]
    t_dev = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    total_ms = float(t_dev.item())
    stats = getattr(strips, "last_stats", None)

    # ---- e2e: the public call with host buffers (pageable, as the Haskell caller has them) --------
    e2e_times = []
    h2d = (scene.geometry.nbytes + scene.substances.nbytes + scene.picture_bytes.nbytes + scene.picture_uses.nbytes +
           strips.entries.nbytes)
    d2h = 4 * scene.width * scene.height
    e2e_pageable = None
    e2e_note = None
    n_e2e = max(3, min(args.steps, 10))
    if world == 1:
        host_img = np.empty((scene.height, scene.width), dtype=np.uint32)
        # the base contract's e2e leg copies from / to pinned host memory: page-lock the caller-side buffers
        # once (gudni_b200_host_register), as a client with long-lived Piles would
        pinned = [a for a in (scene.geometry, scene.substances, scene.entries, scene.picture_bytes, host_img)
                  if a is not None and a.nbytes]
        for a in pinned:
            r.host_register(a)
        # the caller's bitmap is the frame's target (gudni_b200_frame_target_host): the kernels store their rows into it across
        # PCIe while the rest of the frame is rasterized, and frame_end has nothing left to copy
        r.frame_target_host(host_img)
        for i in range(2 + n_e2e):
            barrier()
            t0 = time.perf_counter()
            r.raster_scene(i, scene, out=host_img)
            barrier()
            if i >= 2:
                e2e_times.append(time.perf_counter() - t0)
        direct = host_img.copy()
        r.frame_target_host(None)
        # ... and the same call with the frame rendered in HBM and copied out by frame_end, for comparison and as a check
        tc = []
        for i in range(2 + n_e2e):
            barrier()
            t0 = time.perf_counter()
            r.raster_scene(i, scene, out=host_img)
            barrier()
            if i >= 2:
                tc.append(time.perf_counter() - t0)
        e2e_note = ("frame stored into the caller's page-locked bitmap by the kernels (gudni_b200_frame_target_host); rendered in HBM and "
                    "copied out by frame_end instead: %.1f frames/s; the two bitmaps are equal: %s"
                    % (1.0 / float(np.mean(tc)), bool(np.array_equal(direct, host_img))))
        del direct
        for a in pinned:
            r.host_unregister(a)
        # the same call with the caller's buffers left pageable (what an unmodified Haskell caller has: SURVEY.md §8(b))
        tp = []
        for i in range(2 + n_e2e):
            t0 = time.perf_counter()
            r.raster_scene(i, scene, out=host_img)
            if i >= 2:
                tp.append(time.perf_counter() - t0)
        e2e_pageable = {"value": 1.0 / float(np.mean(tp)), "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "buffers": "pageable host memory (no cudaHostRegister)"}
    else:
        # A host presenter needs no inter-GPU gather at all: every rank rasterizes its strip from host buffers and
        # copies it over ITS OWN PCIe link into its rows of one canvas in POSIX shared memory that every rank has
        # page-locked; the closing barrier is the hand-over to the presenting process.
        from multiprocessing import shared_memory
        name = f"gudni_b200_canvas_{os.environ.get('MASTER_PORT', '0')}"
        shm = None
        if rank == 0:
            try:
                shared_memory.SharedMemory(name=name).unlink()
            except FileNotFoundError:
                pass
            shm = shared_memory.SharedMemory(name=name, create=True, size=d2h)
        barrier()
        if rank != 0:
            shm = shared_memory.SharedMemory(name=name)
            try:   # rank 0 owns the segment: keep this process's resource tracker from unlinking it again at exit
                from multiprocessing import resource_tracker
                resource_tracker.unregister(shm._name, "shared_memory")
            except Exception:  # noqa: BLE001
                pass
        host_canvas = np.ndarray((scene.height, scene.width), dtype=np.uint32, buffer=shm.buf)
        mine = host_canvas[strips.my_rows[0]:strips.my_rows[1]]
        h2d_all = torch.tensor([float(h2d)], dtype=torch.float64, device="cuda")
        dist.all_reduce(h2d_all)
        h2d = int(h2d_all.item())
        pinned = [a for a in (scene.geometry, scene.substances, strips.entries, scene.picture_bytes, mine) if a is not None and a.nbytes]
        for a in pinned:
            r.host_register(a)
        for i in range(2 + n_e2e):
            barrier()
            t0 = time.perf_counter()
            strips.render_to_host(i, host_canvas)
            barrier()
            if i >= 2:
                e2e_times.append(time.perf_counter() - t0)
        for a in pinned:
            r.host_unregister(a)
        e2e_note = ("every rank uploads its inputs and copies its strip over its own PCIe link into one page-locked canvas "
                    "in POSIX shared memory; no NVLink gather for a host presenter")
        if rank == 0:
            e2e_parity = None
            try:
                e2e_parity = bool(np.array_equal(host_canvas, strips.canvas.cpu().numpy().view(np.uint32))) if strips.canvas is not None else None
            except Exception:  # noqa: BLE001
                pass
            e2e_note += f"; canvas equals the device-gathered frame: {e2e_parity}"
        del mine, host_canvas
        barrier()
        shm.close()
        if rank == 0:
            shm.unlink()
    t_e2e = torch.tensor([float(np.mean(e2e_times))], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)

    if rank == 0:
        peak, peak_src = load_peaks()
        ms_per_step = total_ms / args.steps
        value = 1e3 / ms_per_step
        a_bytes = stats.algorithmic_bytes if stats is not None else 0
        k_ms = float(np.mean(raster_ms)) if raster_ms else ms_per_step
        achieved = a_bytes / (k_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get(args.workload)
        line = {
            "metric": "frames/s", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload][1], "canvas": [scene.width, scene.height],
                       "spec": "G=256 MAXT=1024 maxStrandsPerTile=1022 MAXSHAPE=127",
                       "l2": "flushed between timed iterations (256 MiB write)",
                       "parallelism": "1 GPU, whole frame" if world == 1 else
                       f"{world} tile-row strips, gather={(gather_order or {}).get('gather', args.gather)}" + (f", pipelined in chunks of {args.chunk_rows} rows" if pipelined else ""),
                       "strips": strips.rows if world > 1 else None,
                       "gather_order": gather_order},
            "mpixel_per_s": scene.width * scene.height * value / 1e6,
            # SURVEY.md §8(d) secondary work units
            "mthreshold_per_s": (stats.n_thresholds * value / 1e6) if (stats is not None and world == 1) else None,
            "staged_bytes": int(a_bytes + 2 * 20 * stats.n_thresholds) if (stats is not None and world == 1) else None,
            "clocks": clocks,
            "e2e": {"value": 1.0 / float(t_e2e.item()), "unit": "frames/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), **({"note": e2e_note} if e2e_note else {})},
            "e2e_pageable": e2e_pageable,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "the frame's raster kernels, one launch each: raster_generate_kernel, raster_sort_kernel, "
                                   "raster_slice_kernel, raster_resolve_kernel, raster_composite_kernel, raster_accumulate_kernel "
                                   "(+ raster_slice_wide_kernel, raster_picture_kernel, raster_spill_kernel); the dominant one is raster_slice_kernel "
                                   "(profiles/r2_launches_final.csv)", "kernel_ms": k_ms,
                         "bin_ms": float(np.mean(bin_ms)) if bin_ms else None, "algorithmic_bytes": int(a_bytes),
                         "note": "the path is instruction-issue / latency bound, not bandwidth bound: per S4 frame 14.1 M "
                                 "thresholds of curve subdivision, 19 M active runs, 75 M sweep sections, 15.7 M stack composites of "
                                 "~38 layers; ncu per kernel in profiles/r2_final_*.txt"},
            "frame": stats.as_dict() if stats is not None else None,
        }
        if per_rank is not None:
            line["per_rank_raster_bin_ms"] = per_rank
            slowest = max(a + b for a, b in per_rank)
            line["exposed_gather_ms"] = max(0.0, ms_per_step - slowest)   # frame minus the slowest rank's kernels: gather + launch/host gaps
            line["parity_ok"] = parity_ok
            line["parity_note"] = "torch.equal(canvas gathered from N GPUs, the same frame rendered on one GPU), checked on rank 0 before the timed steps"
        if single is not None:
            line["single_gpu_same_workload"] = single
            line["speedup_vs_single_gpu"] = value / single["value"]
        if world == 1 and args.workload == "S4" and not args.no_also:
            # BASELINE.md §3 asks for both readings of "100k": S4 (100k shapes, the value above) and S4b
            # (6,250 circles = exactly 100k curves); same timing method, not part of `value`
            sb = make_scene("S4b")
            sr = StripRenderer(r, sb, 0, 1, None)
            db = DeviceScene(r, sb)
            for i in range(3):
                sr.render(i, db)
            tb = []
            for i in range(10):
                flush.fill_(i & 0xFF)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); sr.render(3 + i, db); e1.record(stream)
                torch.cuda.synchronize()
                tb.append(e0.elapsed_time(e1))
            line["also"] = {"S4b": {"workload": WORKLOADS["S4b"][1], "value": 1e3 / float(np.mean(tb)), "unit": "frames/s",
                                    "ms_per_step": float(np.mean(tb)),
                                    "mpixel_per_s": sb.width * sb.height * 1e3 / float(np.mean(tb)) / 1e6}}
            sr.close(); db.free()
        if world == 1 and args.workload == "S4" and not args.no_also:
            # the multi-GPU workload on ONE GPU, so that the 1 -> 8 curve of the scaling run can be read on one workload
            s5 = make_scene("S5")
            sr = StripRenderer(r, s5, 0, 1, None)
            d5 = DeviceScene(r, s5)
            for i in range(3):
                sr.render(i, d5)
            t5 = []
            for i in range(5):
                flush.fill_(i & 0xFF)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); sr.render(3 + i, d5); e1.record(stream)
                torch.cuda.synchronize()
                t5.append(e0.elapsed_time(e1))
            line["also"]["S5_single_gpu"] = {"workload": WORKLOADS["S5"][1], "value": 1e3 / float(np.mean(t5)), "unit": "frames/s",
                                             "ms_per_step": float(np.mean(t5)),
                                             "mpixel_per_s": s5.width * s5.height * 1e3 / float(np.mean(t5)) / 1e6}
            sr.close(); d5.free(); del sr, s5
        if world == 1 and args.workload == "S4" and not args.no_also:
            try:
                line["also"]["level3_outlines_in"] = level3_reading(r, scene, dscene, stream, flush, torch)
            except Exception as e:  # noqa: BLE001 - a secondary reading never fails the bench
                line["also"]["level3_outlines_in"] = {"unavailable": repr(e)[:300]}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle
            use_all_host_cores()
            kind = cpu_kind()
            run, desc = oracle_frame_sampler(scene, budget_s=10.0, kind=kind)
            t = float(np.mean([run() for _ in range(2)]))
            line["cpu_baseline"] = {"value": 1.0 / t, "unit": "frames/s", "cores": oracle.host_threads(),
                                    "kind": kind, "sample": desc, "note": CPU_NOTES[kind]}
            if kind == "reference":
                # the restated port beside it: same arithmetic, tighter scratch layout and sort
                run, desc = oracle_frame_sampler(scene, budget_s=6.0, kind="port")
                line["cpu_baseline"]["port_value"] = 1.0 / run()
            if not args.no_opencl_reference:
                line["reference_opencl_same_gpu"] = opencl_reference_on_this_gpu(args.workload)
        print(json.dumps(line), flush=True)
    strips.close()
    dscene.free()
    r.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--gather", default="auto", choices=["auto", "nccl", "p2p"],
                    help="N > 1: how strips reach the presenting rank: NCCL send/recv, direct stores over NVLink into its "
                         "CUDA-IPC-mapped canvas, or (auto) whichever the warm-up frames show to be faster")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary S4b reading")
    ap.add_argument("--no-opencl-reference", action="store_true",
                    help="skip the side measurement of the reference's kernels under OpenCL on this GPU")
    ap.add_argument("--no-rebalance", action="store_true", help="N > 1: keep the area-based strip partition")
    ap.add_argument("--gather-order", choices=("auto", "early", "late"), default="auto",
                    help="N > 1: post the presenting rank's receives before (early) or after (late) its own strip; auto measures both")
    ap.add_argument("--rebalance-rounds", type=int, default=4, help="N > 1: feedback rounds of the strip rebalancer")
    ap.add_argument("--gather-gbs", type=float, default=800.0,
                    help="N > 1: inbound GB/s of the presenting rank assumed by the strip rebalancer (measured ~780 on NVLink 5)")
    ap.add_argument("--pipeline", action="store_true",
                    help="N > 1: send the strip chunk by chunk while rendering the next chunk (measured slower on S5: "
                         "a 256-row chunk cannot fill a B200, see DESIGN.md §5)")
    ap.add_argument("--chunk-rows", type=int, default=512, help="N > 1: rows per pipelined chunk (multiple of 256)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is None:
        # the metric is quoted on the 4K scene; a frame that fits one GPU is not sharded
        # (BASELINE.json north_star), so N > 1 runs the 16K^2 canvas the strips exist for
        args.workload = "S4" if max(world, args.gpus) == 1 else "S5"
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_native(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
