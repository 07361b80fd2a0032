#!/usr/bin/env python
"""Benchmark of the rasteriser hot path (BASELINE.json metric: fwd+bwd iterations/s and forward
FPS, 3 M Gaussians, SH degree 3, 1297x840).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA kernels via the C ABI)
    python bench.py --impl reference --gpus N ...            # the reference's own implementation

A "step" is one pass of the hot path over one view: SHRenderer.forward (cull -> project -> bin/sort
-> composite) + L2 loss against a target image + backward to the leaf parameters.  Workload at
N = 1 is cfg 2 of BASELINE.md.  At N > 1 every rank renders one view of the same replicated scene
per step (views a few degrees apart) and the parameter gradients (236 B/Gaussian) are summed with
one NCCL all-reduce: weak scaling, value = views/s over all ranks.

JSON keys beyond the base contract: `fwd_fps` (forward-only FPS), `kernels_ms` (CUDA-event time
of each stage), `roofline` (dominant kernel), `cpu_baseline`, `e2e`, `clocks`, `gpu_launches`.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg5"])
    ap.add_argument("--n-gaussians", type=int, default=None, help="override N (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exact", action="store_true", help="disable exact skip decisions")
    ap.add_argument("--dp-mode", default="auto", choices=["auto", "pull", "push", "fused", "sparse", "allreduce"],
                    help="N>1: 'fused' = SH gradients reduced into all ranks by the backward kernel over "
                         "NVLink/NVSwitch + small all-reduce; 'sparse' = all-reduce of the union of touched "
                         "rows only; 'allreduce' = one dense NCCL all-reduce; 'push' (= 'auto') = each rank "
                         "adds the rows it touched into every rank's result buffer over NVSwitch multicast")
    return ap.parse_args()


# ---------------------------------------------------------------- clocks

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([x.strip() for x in r.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def phys_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        parts = [p for p in vis.split(",") if p != ""]
        if local_rank < len(parts) and parts[local_rank].isdigit():
            return int(parts[local_rank])
    return local_rank


# ---------------------------------------------------------------- shared helpers

def views_for(world, torch):
    """Identity pose for rank 0's cfg-2 view; other ranks yaw by 2 degrees steps around it."""
    import math

    out = []
    for i in range(world):
        a = math.radians(2.0 * (i - (world - 1) / 2.0)) if world > 1 else 0.0
        c2w = torch.tensor([[math.cos(a), 0.0, math.sin(a), 0.0],
                            [0.0, 1.0, 0.0, 0.0],
                            [-math.sin(a), 0.0, math.cos(a), 0.0]], dtype=torch.float32)
        out.append(c2w)
    return out


def dp_exchange_desc(flat, world, N):
    if world == 1 or flat is None:
        return None
    if flat.pull:
        how = ("multimem.ld_reduce (summed in the NVSwitch) + multimem.st" if flat.multicast_ptr
               else "peer loads + peer stores over NVLink")
        return f"sparse all-reduce of the union of touched rows (240 B each) over symmetric memory: {how}"
    if flat.push:
        how = ("multimem.red on the NVSwitch multicast address" if flat.multicast_ptr
               else "red.global.add per peer over NVLink")
        return f"each rank pushes the rows it touched (240 B each) into every rank's result buffer: {how}"
    if flat.fused:
        return "in-kernel multimem/peer reduction of SH grads + 132 MB all-reduce"
    if flat.sparse:
        return (f"NCCL all-reduce of the union of touched rows ({flat.last_union_rows} of {N} rows x 240 B) "
                "+ 3 MB mark all-reduce")
    return "dense 708 MB NCCL all-reduce"


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(args, scene_name, seed=0):
    """The oracle (C restatement of the reference kernels + the reference's torch-level ops, CPU)
    timed on a bounded sample: the first n_s Gaussians of the workload's scene, full image, one
    forward+backward; scaled to the metric's unit by N / n_s (duplicate work is linear in N)."""
    import torch

    from gaussian_splatting_3d_b200 import synthetic as S
    from oracle import ref_torch as R

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    N_full, C = S.CONFIGS[scene_name][0], S.CONFIGS[scene_name][1]
    n_s = min(N_full, 100_000)
    cam = S.make_camera(scene_name)
    sc = S.make_scene(scene_name, seed=seed, N=n_s)
    names = ("mean", "qvec", "svec_before_activation", "sh_coeffs", "alpha_before_activation")
    p = {k: sc[k].clone().requires_grad_(True) for k in names}
    tgt = S.make_target(cam, seed)
    t0 = time.perf_counter()
    img = R.reference_forward(p, sc["c2w"], cam, C)
    t1 = time.perf_counter()
    ((img - tgt) ** 2).mean().backward()
    t2 = time.perf_counter()
    scale = N_full / n_s
    return {"value": 1.0 / ((t2 - t0) * scale), "unit": "iters/s", "cores": cores, "kind": "port",
            "sample": f"first {n_s} of {N_full} Gaussians of {scene_name} (seed {seed}), full {cam.w}x{cam.h} image, "
                      f"1 fwd+bwd = {t2 - t0:.2f} s (fwd {t1 - t0:.2f} s), scaled by N/n_s = {scale:.0f}",
            "fwd_fps": 1.0 / ((t1 - t0) * scale)}


# ---------------------------------------------------------------- our arm

def run_ours(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist

    from gaussian_splatting_3d_b200 import capi, ops
    from gaussian_splatting_3d_b200 import synthetic as S

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    name = args.workload
    cam = S.make_camera(name)
    sc = S.make_scene(name, seed=0, N=args.n_gaussians)
    C = sc["C"]
    cfg = S.make_cfg(device=str(dev), sh_order=C, exact_decisions=not args.no_exact)
    r = S.renderer_from_scene(sc, cfg)
    r.train()
    N = r.N
    c2w_host = views_for(world, torch)[rank].pin_memory()
    tgt_host = S.make_target(cam, rank).pin_memory()
    c2w_dev = c2w_host.to(dev)
    tgt_dev = tgt_host.to(dev)
    params = [r.mean, r.qvec, r.svec_before_activation, r.sh_coeffs, r.alpha_before_activation]
    flat = None
    if world == 1:
        # same flat-gradient plumbing as the multi-GPU path, minus the exchange: the backward kernels add
        # into one persistent buffer aliased by the parameters' .grad, and the per-step reset clears only
        # the rows the previous backward touched (no 708 MB fill per step)
        from gaussian_splatting_3d_b200 import parallel as P

        flat = P.FlatGradients(r, sparse_reset=True).attach(r)
    if world > 1:
        # one flat gradient buffer that the backward kernels write into directly (no packing copy),
        # summed over the ranks with a single NCCL all-reduce per step
        from gaussian_splatting_3d_b200 import parallel as P

        mode = args.dp_mode if args.dp_mode != "auto" else "pull"
        flat = P.FlatGradients(r, fused=(mode == "fused"), sparse=(mode == "sparse"), push=(mode == "push"),
                               pull=(mode == "pull")).attach(r)

    copy_stream = torch.cuda.Stream(device=dev)

    def step(e2e):
        tgt_ready = None
        if e2e:
            # the pose (48 B) is needed by the first kernel: current stream.  The 13 MB target image is not
            # needed before the loss: its host->device copy runs on a copy stream beside the forward kernels
            c2w = c2w_host.to(dev, non_blocking=True)
            main = torch.cuda.current_stream(dev)
            copy_stream.wait_stream(main)  # (the previous step's consumers of the recycled block are done)
            with torch.cuda.stream(copy_stream):
                tgt = tgt_host.to(dev, non_blocking=True)
                tgt_ready = torch.cuda.Event()
                tgt_ready.record(copy_stream)
            tgt.record_stream(main)
        else:
            c2w, tgt = c2w_dev, tgt_dev
        flat.zero()  # per-step reset of the gradient buffers (N > 1: side stream, overlaps the forward)
        out = r(c2w, cam)
        if tgt_ready is not None:
            torch.cuda.current_stream(dev).wait_event(tgt_ready)
        loss = ((out - tgt) ** 2).mean()
        flat.backward_into(loss)
        if world > 1:
            flat.exchange()
        if e2e:
            return float(loss.item())  # device -> host read of the step's result
        return loss

    def fwd_only():
        with torch.no_grad():
            return r(c2w_dev, cam)

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(max(args.warmup, 3)):
        step(True)
        step(False)
        fwd_only()
    torch.cuda.synchronize()

    with ClockSampler(phys_gpu_index(local_rank)) as clk:
        l0 = capi.lib.gs3d_launch_count()
        ms_total = timed(lambda: step(False), args.steps)
        launches = capi.lib.gs3d_launch_count() - l0
        ms_e2e = timed(lambda: step(True), args.steps)
        ms_fwd = timed(fwd_only, args.steps)
    clocks = clk.summary()

    # ---- per-stage CUDA-event times (same stream the kernels are launched on)
    stage_ms = {}
    orig = {}

    def wrap(fname, label):
        f = getattr(ops, fname)
        orig[fname] = f

        def g(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = f(*a, **k)
            e1.record()
            stage_ms.setdefault(label, []).append((e0, e1))
            return out

        setattr(ops, fname, g)

    for fname, label in (("project_cull_fused", "K1_project_cull"), ("tile_culling_aabb_start_end", "K2_binning"),
                         ("composite_sh_forward", "K3_composite_fwd"), ("composite_sh_backward", "K4a_composite_bwd"),
                         ("project_backward_fused", "K4b_project_bwd"), ("rows_push_marked", "X_rows_push"),
                         ("rows_pull_marked", "X_rows_pull"), ("marks_broadcast", "X_marks_broadcast"),
                         ("rows_zero_marked", "X_rows_zero")):
        wrap(fname, label)
    if flat is not None and world > 1:  # the data-parallel exchange: per-step reset (+ barrier) and exchange (+ barrier)
        for meth, label in (("zero", "DP_reset"), ("exchange", "DP_exchange")):
            def timed_method(f=getattr(flat, meth), label=label):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = f()
                e1.record()
                stage_ms.setdefault(label, []).append((e0, e1))
                return out
            setattr(flat, meth, timed_method)
    n_prof = min(args.steps, 10)
    # units the compositing launches really process: duplicates STAGED into shared memory (tiles stop
    # staging once every pixel is saturated), counted by the kernels themselves during these steps
    staged = torch.zeros(2, dtype=torch.int64, device=dev)
    ops.set_stage_counters(staged)
    for _ in range(n_prof):
        step(False)
    torch.cuda.synchronize()
    ops.set_stage_counters(None)
    staged_fwd, staged_bwd = (int(x) // n_prof for x in staged.tolist())
    for fname, f in orig.items():
        setattr(ops, fname, f)
    if flat is not None and world > 1:
        del flat.zero, flat.exchange  # drop the instance-level wrappers
    kernels_ms = {k: sum(a.elapsed_time(b) for a, b in v) / n_prof for k, v in stage_ms.items()}  # per step

    rank_kernel_ms = None
    if world > 1:  # per-rank sum of the hot-path kernels: shows how uneven the views' work is
        mine = torch.tensor([sum(v for k, v in kernels_ms.items() if k.startswith("K"))], device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_kernel_ms = [round(float(t.item()), 4) for t in allr]
    n_dub = r.total_dub_gaussians
    ms_step = ms_total / args.steps
    value = world * 1000.0 / ms_step
    hbm_peak, peak_src = peaks()
    CC = C * C
    # algorithmic bytes per launch of the dominant kernel (DESIGN.md "Kernels"): per STAGED
    # duplicate 4 (id) + 48 (record) + 12*C^2 (SH row); per pixel 12 (image) [+ 24 read in backward];
    # backward adds one gradient row reduction of 4*(7+3C^2) B per staged duplicate at most (rows of
    # Gaussians that contributed nothing are skipped).  `upper_bound` is the same with all n_dub staged.
    px = cam.w * cam.h
    per_dup = 4 + 48 + 12 * CC
    bytes_fwd = staged_fwd * per_dup + px * 12
    bytes_bwd = staged_bwd * (per_dup + 4 * (7 + 3 * CC)) + px * 36
    upper = {"K3_composite_fwd": n_dub * per_dup + px * 12,
             "K4a_composite_bwd": n_dub * (per_dup + 4 * (7 + 3 * CC)) + px * 36}
    kk = {k: v for k, v in kernels_ms.items() if k.startswith("K")}
    dom = max(kk, key=kk.get) if kk else None
    algo = {"K3_composite_fwd": bytes_fwd, "K4a_composite_bwd": bytes_bwd, "K1_project_cull": N * (48 + 101),
            "K2_binning": N * (4 + 4 * 24 + 8) + n_dub * (8 + 2 * 24 + 4), "K4b_project_bwd": N * (48 + 28 + 44 + 1)}
    roofline = None
    if dom:
        ach = algo[dom] / (kernels_ms[dom] * 1e-3) / 1e9
        traffic, traffic_src, issue = None, None, None
        tfile = ROOT / "profiles" / "ncu_traffic.json"
        if name == "cfg2" and args.n_gaussians is None and tfile.exists():
            try:  # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this workload
                t = json.loads(tfile.read_text()).get(dom)
                if t:
                    traffic, traffic_src = t["traffic"], t["source"]
                    # what actually bounds the compositing kernels (same capture): issue slots / pipes
                    issue = {k: t[k] for k in ("issue_slot_pct", "fma_pipe_pct", "lsu_pipe_pct", "smem_wavefront_pct",
                                               "warps_active_pct") if t.get(k) is not None}
            except Exception:
                pass
        roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": traffic, "traffic_source": traffic_src,
                    "peak_source": peak_src, "algorithmic_bytes": algo[dom], "launch_ms": kernels_ms[dom],
                    "staged_duplicates": {"fwd": staged_fwd, "bwd": staged_bwd, "n_dub": n_dub},
                    "ncu_utilisation_pct_of_peak": issue,
                    "algorithmic_bytes_if_all_staged": upper.get(dom),
                    "note": "compositing is FP32-issue / shared-memory bound, not HBM bound (DESIGN.md: ncu "
                            "issue-slot utilisation 64-69 %, FMA pipe 42-47 %, DRAM < 2 %); algorithmic bytes "
                            "count the duplicates actually staged (tiles stop once saturated) and are served "
                            "mostly from L2 (a Gaussian is staged by ~3.7 tiles), hence traffic < algorithmic"}
    line = {
        "metric": "fwd+bwd iters/s (3M Gaussians SH3 @1297x840)", "value": value, "unit": "iters/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name}: {N} Gaussians, SH degree {C - 1} (C={C}), {cam.w}x{cam.h}, "
                               f"1 view/step/GPU, fwd + L2 loss + bwd", "n_dub": n_dub,
                   "views_per_step": world, "parallelism": f"dp{world}" if world > 1 else "single",
                   "dp_exchange": dp_exchange_desc(flat, world, N),
                   "gradient_buffers": "one persistent flat buffer aliased by .grad; per-step reset clears only "
                                       "the rows the previous backward marked",
                   "l2_policy": "inputs larger than L2 (parameters 708 MB, duplicates 132 MB vs 126 MB L2)",
                   "exact_decisions": not args.no_exact},
        "fwd_fps": world * 1000.0 * args.steps / ms_fwd,
        "kernels_ms": kernels_ms,
        "rank_kernel_ms": rank_kernel_ms,
        "roofline": roofline,
        "e2e": {"value": world * 1000.0 * args.steps / ms_e2e, "unit": "iters/s",
                "h2d_bytes_per_step": int(c2w_host.numel() * 4 + tgt_host.numel() * 4), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args, name)
            except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"error": str(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------- reference arm

def run_reference(args, rank, local_rank, world):
    """The reference's own implementation: its CUDA extension (oracle/_ref, built from the
    unmodified sources + the 6-line scratch patch that makes HEAD compile) behind its own torch-level
    flow, on the same scene / camera / loss.  Rank 0 only.  When the extension is not available the
    CPU oracle port is timed instead (bounded sample)."""
    if rank != 0:
        return
    import torch

    from gaussian_splatting_3d_b200 import synthetic as S

    name = args.workload
    cam = S.make_camera(name)
    cpu = None
    if not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(args, name)
        except Exception as e:
            cpu = {"error": str(e)}
    ext = None
    why = ""
    if torch.cuda.is_available():
        try:
            from oracle import ref_gpu

            ext = ref_gpu.load_reference_extension()
            if ext is None:
                why = "oracle/_ref/_gs_ref*.so absent"
        except Exception as e:
            why = f"reference extension not loadable: {e}"
    else:
        why = "no CUDA device"
    base = {"impl": "reference", "metric": "fwd+bwd iters/s (3M Gaussians SH3 @1297x840)", "unit": "iters/s",
            "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "cpu_baseline": cpu}
    if ext is None:
        v = cpu["value"] if cpu and "value" in cpu else None
        base.update({"value": v, "ms_per_step": (1000.0 / v) if v else None,
                     "config": {"workload": f"{name} on host cores (oracle port; {why})"},
                     "e2e": {"value": v, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(base), flush=True)
        return
    from oracle import ref_gpu

    dev = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(dev)
    sc = S.make_scene(name, seed=0, N=args.n_gaussians)
    C = sc["C"]
    ref = ref_gpu.ReferenceGPURenderer(ext, sc, dev, C)
    c2w_host = sc["c2w"].pin_memory()
    tgt_host = S.make_target(cam, 0).pin_memory()
    c2w_dev, tgt_dev = c2w_host.to(dev), tgt_host.to(dev)

    def step(e2e):
        c2w = c2w_host.to(dev, non_blocking=True) if e2e else c2w_dev
        tgt = tgt_host.to(dev, non_blocking=True) if e2e else tgt_dev
        ref.zero_grad()
        out = ref.forward(c2w, cam)
        loss = ((out - tgt) ** 2).mean()
        loss.backward()
        return float(loss.item()) if e2e else loss

    def fwd_only():
        with torch.no_grad():
            return ref.forward(c2w_dev, cam)

    def timed(fn, k):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    # The reference's backward printf()s from the device whenever its recomputed image differs
    # from the saved one (vol_render_sh.h:448-451); keep that out of our single JSON line.
    sys.stdout.flush()
    saved_fd = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        for _ in range(max(args.warmup, 3)):
            step(True)
        with ClockSampler(phys_gpu_index(local_rank)) as clk:
            ms = timed(lambda: step(False), args.steps)
            ms_e2e = timed(lambda: step(True), args.steps)
            ms_fwd = timed(fwd_only, args.steps)
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(devnull)
        os.close(saved_fd)
    base.update({
        "value": 1000.0 * args.steps / ms, "ms_per_step": ms / args.steps,
        "fwd_fps": 1000.0 * args.steps / ms_fwd,
        "config": {"workload": f"{name}: {ref.params['mean'].shape[0]} Gaussians, C={C}, {cam.w}x{cam.h}, reference "
                               "CUDA extension (sm_100a build of /root/reference/gs/src, -DNDEBUG) behind the "
                               "reference's torch-level flow", "n_dub": ref.total_dub_gaussians},
        "e2e": {"value": 1000.0 * args.steps / ms_e2e, "unit": "iters/s",
                "h2d_bytes_per_step": int(c2w_host.numel() * 4 + tgt_host.numel() * 4), "d2h_bytes_per_step": 4},
        "clocks": clk.summary(),
        "reference_kind": "gpu-extension",
    })
    print(json.dumps(base), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if args.impl == "reference":
        run_reference(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
