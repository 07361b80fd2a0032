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
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="cfg2 (default): one view per rank per step, weak scaling.  cfg4 (SURVEY 8d): the cfg-2 scene, "
                         "8 ring cameras per step, 8/G cameras per rank, gradients exchanged -- strong scaling.  cfg5: "
                         "6 M Gaussians at 3840x2160, forward only, tile rows sharded over the ranks -- strong scaling")
    ap.add_argument("--no-graph", action="store_true",
                    help="time the eager step (one 8-byte read-back of the duplicate count per forward, ~40 launches) "
                         "instead of the CUDA-graph step (graph.GraphedStep; single-GPU single-view workloads)")
    ap.add_argument("--check", action="store_true",
                    help="after timing: whole-path parity of this workload against the reference GPU flow "
                         "(oracle/fullsize_check.py; needs oracle/_ref) added to the line as `parity`")
    ap.add_argument("--n-gaussians", type=int, default=None, help="override N (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-exact", action="store_true", help="disable exact skip decisions")
    ap.add_argument("--dp-mode", default="auto", choices=["auto", "pull", "push", "fused", "sparse", "allreduce"],
                    help="N>1: 'fused' = SH gradients reduced into all ranks by the backward kernel over "
                         "NVLink/NVSwitch + small all-reduce; 'sparse' = all-reduce of the union of touched "
                         "rows only; 'allreduce' = one dense NCCL all-reduce; 'push' (= 'auto') = each rank "
                         "adds the rows it touched into every rank's result buffer over NVSwitch multicast")
    ap.add_argument("--band-gather", default="root", choices=["root", "allgather"],
                    help="cfg5, N>1: 'root' = bands stored straight into rank 0's frame over NVLink (symmetric "
                         "memory); 'allgather' = every rank receives the full frame through one padded all_gather")
    ap.add_argument("--yaw-step", type=float, default=0.5,
                    help="N>1 (weak scaling, one view per rank): yaw between neighbouring ranks' poses in degrees.  "
                         "The cfg-2 scene covers the image with a 5 %% margin (about +-3.8 degrees of yaw): inside "
                         "it every rank's view is the cfg-2 workload; further out (e.g. the 2-degree steps of round "
                         "1 at 8 ranks: +-7 degrees) the scene's lateral boundary enters the frame, its tiles never "
                         "saturate and those views cost up to 13 %% more on ONE GPU (tools/view_costs.py)")
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

def views_for(world, torch, yaw_step=0.5):
    """One pose per rank: the cfg-2 identity pose yawed by `yaw_step` degrees between neighbouring ranks,
    centred on the identity (world == 1: the identity itself)."""
    import math

    out = []
    for i in range(world):
        a = math.radians(yaw_step * (i - (world - 1) / 2.0)) if world > 1 else 0.0
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


def workload_desc(name, N, C, cam, views_per_step):
    """The same string in both arms (the driver compares `config` of the two lines)."""
    if name == "cfg5":
        return f"cfg5: {N} Gaussians, SH degree {C - 1} (C={C}), {cam.w}x{cam.h}, forward only"
    what = "8 ring cameras per step (radius 7 around (0,0,7))" if name == "cfg4" else "1 view/step/GPU"
    return f"{name}: {N} Gaussians, SH degree {C - 1} (C={C}), {cam.w}x{cam.h}, {what}, fwd + L2 loss + bwd"


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


# ---------------------------------------------------------------- multi-GPU checks and side workloads

def exchange_check(flat, step_once, torch, dist):
    """Run ONE more step and compare the exchanged gradient buffer with a dense NCCL all-reduce of the ranks'
    own contributions of that same step -> max over ranks of ||exchanged - dense|| / ||dense||."""
    own = {}
    orig = flat.exchange

    def spy(*a, **k):
        # this rank's contribution: the private rows (push / pull), else the not yet reduced flat buffer
        src = flat.local if flat.local is not None else flat.flat
        own["g"] = src.detach().clone()
        return orig(*a, **k)

    if flat.fused:
        return None  # the SH block is reduced inside the backward kernel: no separate own contribution exists
    flat.exchange = spy
    try:
        step_once()
    finally:
        del flat.exchange
    torch.cuda.synchronize()
    dense = own["g"]
    dist.all_reduce(dense, op=dist.ReduceOp.SUM)
    num = (flat.flat.double() - dense.double()).norm()
    den = dense.double().norm().clamp_min(1e-30)
    rel = (num / den).float().reshape(1)
    dist.all_reduce(rel, op=dist.ReduceOp.MAX)
    return float(rel.item())


def union_rows_of_last_exchange(flat, torch):
    """Rows (Gaussians) whose gradients crossed NVLink in the last exchange, and the bytes each GPU ingested."""
    if flat is None or not flat.push:
        return None, None
    used = flat._res[flat._cur ^ 1]  # exchange() flipped the buffers
    U = int((used["union"] != 0).sum().item())
    row_bytes = 4 * sum(v.numel() // v.size(0) for v in flat.views)
    return U, U * row_bytes + int(used["union"].numel())


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
    if name == "cfg5":
        return run_cfg5(args, rank, local_rank, world, dev)
    cfg4 = name == "cfg4"
    scene = "cfg2" if cfg4 else name  # cfg 4 = the cfg-2 scene seen from 8 ring cameras
    cam = S.make_camera(scene)
    sc = S.make_scene(scene, seed=0, N=args.n_gaussians)
    C = sc["C"]
    cfg = S.make_cfg(device=str(dev), sh_order=C, exact_decisions=not args.no_exact)
    r = S.renderer_from_scene(sc, cfg)
    r.train()
    N = r.N
    if cfg4:
        from gaussian_splatting_3d_b200 import parallel as P

        ring = S.ring_cameras(8)
        my_views = P.shard_views(8, rank, world)  # 8/G cameras per rank, round-robin
        c2w_hosts = [ring[i].pin_memory() for i in my_views]
        tgt_hosts = [S.make_target(cam, i).pin_memory() for i in my_views]
        n_views_step = 8
    else:
        c2w_hosts = [views_for(world, torch, args.yaw_step)[rank].pin_memory()]
        tgt_hosts = [S.make_target(cam, rank).pin_memory()]
        n_views_step = world
    c2w_devs = [c.to(dev) for c in c2w_hosts]
    tgt_devs = [t.to(dev) for t in tgt_hosts]
    c2w_dev = c2w_devs[0]
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
    c2w_stage = [torch.empty_like(c) for c in c2w_devs]  # device staging buffers of the end-to-end path
    tgt_stage = [torch.empty_like(t) for t in tgt_devs]
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    loss_ready = torch.cuda.Event()

    def step(e2e):
        flat.zero()  # per-step reset of the gradient buffers (N > 1: side stream, overlaps the forward)
        total = None
        for v in range(len(c2w_devs)):
            tgt_ready = None
            if e2e:
                # the pose (48 B) is needed by the first kernel: current stream.  The 13 MB target image is not
                # needed before the loss: its host->device copy runs on a copy stream beside the forward kernels.
                # Both land in persistent staging buffers (no allocation, no record_stream bookkeeping per step).
                c2w = c2w_stage[v]
                c2w.copy_(c2w_hosts[v], non_blocking=True)
                main = torch.cuda.current_stream(dev)
                copy_stream.wait_stream(main)  # (the previous step's loss has read this staging buffer)
                with torch.cuda.stream(copy_stream):
                    tgt = tgt_stage[v]
                    tgt.copy_(tgt_hosts[v], non_blocking=True)
                    tgt_ready = torch.cuda.Event()
                    tgt_ready.record(copy_stream)
            else:
                c2w, tgt = c2w_devs[v], tgt_devs[v]
            out = r(c2w, cam)
            if tgt_ready is not None:
                torch.cuda.current_stream(dev).wait_event(tgt_ready)
            loss = ((out - tgt) ** 2).mean()
            total = loss.detach() if total is None else total + loss.detach()
            if e2e and len(c2w_devs) == 1:
                # one view per step: the step's result is final once the loss exists: its 4-byte device -> host copy
                # is enqueued BEFORE the backward (and the exchange), so the host has it while they still run
                loss_host.copy_(total.reshape(1), non_blocking=True)
                loss_ready.record(torch.cuda.current_stream(dev))
            flat.backward_into(loss)  # accumulates into the flat buffer
        if world > 1:
            flat.exchange()
            if cfg4:  # a training step: the ADC statistics follow the gradients (parallel.view_sharded_step)
                P.sync_adc(r)
        if e2e:
            if len(c2w_devs) == 1:
                loss_ready.synchronize()
                return float(loss_host[0])  # device -> host read of the step's result
            # several views per step: read at the end of the step (letting the host run a whole multi-view eager
            # step ahead makes the caching allocator grow instead of recycling: cfg 4 on one GPU 316 -> 187 it/s)
            return float(total.item())
        return total

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
    sync_free_capacity = None
    if (world > 1 or cfg4) and not args.no_graph:
        # multi-view / multi-GPU steps run eagerly (the exchange alternates between two symmetric buffers), but
        # without the per-forward host read-back of the duplicate count: static id capacity = 1.25 x the largest
        # count among this rank's views (SHRenderer.static_capacity; the overflow flag is checked after the timing)
        worst = 0
        with torch.no_grad():
            for c in c2w_devs:
                r(c, cam)
                worst = max(worst, int(r.total_dub_gaussians))
        sync_free_capacity = max(1 << 16, (int(worst * 1.25) + 65535) // 65536 * 65536)
        r.static_capacity = sync_free_capacity
        for _ in range(2):
            step(False)
        torch.cuda.synchronize()
        assert not r.overflowed(), "static capacity too small"

    # single-GPU, one view per step: the product's fast path is the CUDA-graph step (static duplicate capacity, no
    # host round trip, two graph launches per step); `--no-graph` times the eager step instead
    # forward-only FPS: the eager render call (viewer / evaluation loop), one 8-byte read-back of the duplicate count
    # per frame.  (Measured: the same loop with the static capacity and NO per-frame synchronisation is slower, 637
    # against 893 FPS -- the host runs frames ahead and the caching allocator grows instead of recycling.)
    cap_saved, r.static_capacity = r.static_capacity, None
    ms_fwd = timed(fwd_only, args.steps)
    r.static_capacity = cap_saved
    gstep = None
    if world == 1 and not cfg4 and not args.no_graph:
        from gaussian_splatting_3d_b200.graph import GraphedStep

        gstep = GraphedStep(r, flat, cam)
        gstep(c2w_devs[0], tgt_devs[0])  # capture
        gstep.target.copy_(tgt_devs[0])
        gstep.check()
        torch.cuda.synchronize()
    l0 = capi.lib.gs3d_launch_count()
    step(False)
    launches_per_step = capi.lib.gs3d_launch_count() - l0
    with ClockSampler(phys_gpu_index(local_rank)) as clk:
        if gstep is not None:
            r.static_capacity = gstep.capacity
            ms_total = timed(lambda: gstep(c2w_devs[0], gstep.target), args.steps)
            ms_e2e = timed(lambda: gstep(c2w_hosts[0], tgt_hosts[0], read_loss=True), args.steps)
            gstep.check()
        else:
            ms_total = timed(lambda: step(False), args.steps)
            ms_e2e = timed(lambda: step(True), args.steps)
            if sync_free_capacity is not None:
                assert not r.overflowed(), "static capacity overflowed during the timed steps"
        launches = launches_per_step * args.steps  # kernels executed in the timed region (graph nodes included)
    clocks = clk.summary()
    for _ in range(2):  # (graph capture emptied the allocator cache: refill it before the per-stage event timing)
        step(False)

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

    # (the per-stage steps keep the static duplicate capacity of the timed steps: no host read-back inside K1's
    # event pair, same kernels as the timed region)
    for fname, label in (("project_cull_fused", "K1_project_cull"), ("tile_culling_aabb_start_end", "K2_binning"),
                         ("tile_culling_aabb_start_end_capacity", "K2_binning"),
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
    for _ in range(n_prof):  # per-stage times: the production kernels
        step(False)
    torch.cuda.synchronize()
    # unit counts (separate steps: launches made while the counters are set run the instrumented kernels)
    staged = torch.zeros(4, dtype=torch.int64, device=dev)
    ops.set_stage_counters(staged)
    n_cnt = 2
    for _ in range(n_cnt):
        step(False)
    torch.cuda.synchronize()
    ops.set_stage_counters(None)
    staged_fwd, staged_bwd, pairs_fwd, pairs_bwd = (int(x) // n_cnt for x in staged.tolist())
    for fname, f in orig.items():
        setattr(ops, fname, f)
    if flat is not None and world > 1:
        del flat.zero, flat.exchange  # drop the instance-level wrappers
    # per step, over the first n_prof (un-instrumented) steps
    per_step = {k: len(v) // (n_prof + n_cnt) for k, v in stage_ms.items()}
    kernels_ms = {k: sum(a.elapsed_time(b) for a, b in v[:per_step[k] * n_prof]) / n_prof for k, v in stage_ms.items()}

    rank_kernel_ms = None
    if world > 1:  # per-rank sum of the hot-path kernels: shows how uneven the views' work is
        mine = torch.tensor([sum(v for k, v in kernels_ms.items() if k.startswith("K"))], device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        rank_kernel_ms = [round(float(t.item()), 4) for t in allr]
    n_dub = r.total_dub_gaussians
    ms_step = ms_total / args.steps
    value = n_views_step * 1000.0 / ms_step
    rooflines, dom = stage_rooflines(kernels_ms, N=N, C=C, cam=cam, n_dub=n_dub, views=len(c2w_devs),
                                     staged=(staged_fwd, staged_bwd), pairs=(pairs_fwd, pairs_bwd),
                                     ncu_ok=(name == "cfg2" and args.n_gaussians is None),
                                     touched_rows=(int((flat.touched != 0).sum().item())
                                                   if flat is not None and flat.touched is not None else 0))
    n_views_rank = len(c2w_devs)
    h2d = sum(int(c.numel() * 4 + t.numel() * 4) for c, t in zip(c2w_hosts, tgt_hosts))
    line = {
        "metric": "fwd+bwd iters/s (3M Gaussians SH3 @1297x840)", "value": value, "unit": "iters/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong" if cfg4 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_desc(name, N, C, cam, n_views_step), "n_dub": n_dub,
                   "views_per_step": n_views_step, "parallelism": f"dp{world}" if world > 1 else "single",
                   **({"rank_poses": f"cfg-2 pose yawed by {args.yaw_step} degrees between neighbouring ranks"}
                      if (world > 1 and not cfg4) else {})},
        "impl_details": {"views_per_rank": n_views_rank, "dp_exchange": dp_exchange_desc(flat, world, N),
                         "gradient_buffers": "one persistent flat buffer aliased by .grad; per-step reset clears "
                                             "only the rows the previous backward marked",
                         "l2_policy": "inputs larger than L2 (parameters 708 MB, duplicates 132 MB vs 126 MB L2)",
                         "exact_decisions": not args.no_exact,
                         "step_launch": ("3 CUDA-graph launches per step ([reset + forward], [loss], [backward]; graph.GraphedStep: static capacity "
                                         f"{gstep.capacity} duplicates, no host round trip)" if gstep is not None
                                         else ("eager launches, static duplicate capacity "
                                               f"{sync_free_capacity} (no host round trip in the step)"
                                               if sync_free_capacity is not None else
                                               "eager: one 8-byte read-back of the duplicate count per forward")),
                         "kernels_per_step": int(launches_per_step)},
        "fwd_fps": world * 1000.0 * args.steps / ms_fwd,
        "kernels_ms": kernels_ms,
        "rank_kernel_ms": rank_kernel_ms,
        "roofline": rooflines.get(dom) if dom else None,
        "rooflines": rooflines,
        "e2e": {"value": n_views_step * 1000.0 * args.steps / ms_e2e, "unit": "iters/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if world > 1:
        # correctness of the exchange, measured in this very run: one more step against a dense NCCL all-reduce
        line["exchange_check"] = {"rel_err_vs_dense_allreduce": exchange_check(flat, lambda: step(False), torch, dist),
                                  "tolerance": 1e-5}
        U, ingest = union_rows_of_last_exchange(flat, torch)
        line["exchange_union_rows"] = U
        line["exchange_bytes_ingested_per_gpu"] = ingest
        line["exchange_dense_bytes"] = int(flat.flat.numel() * 4)
    if args.check and world == 1:
        line["parity"] = parity_block(name, args, torch)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args, scene)
            except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
                line["cpu_baseline"] = {"error": str(e)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_cfg5(args, rank, local_rank, world, dev):
    """cfg 5 (BASELINE configs[4]): 6 M Gaussians at 3840x2160, forward only, tile rows sharded over the ranks
    (parallel.tile_sharded_render: replicated projection, per-rank band binning + compositing, bands gathered).
    Strong scaling: value = frames/s of the SAME frame at N GPUs."""
    import torch
    import torch.distributed as dist

    from gaussian_splatting_3d_b200 import capi
    from gaussian_splatting_3d_b200 import parallel as P
    from gaussian_splatting_3d_b200 import synthetic as S

    name = "cfg5"
    cam = S.make_camera(name)
    sc = S.make_scene(name, seed=0, N=args.n_gaussians)
    C = sc["C"]
    r = S.renderer_from_scene(sc, S.make_cfg(device=str(dev), sh_order=C, exact_decisions=not args.no_exact))
    r.eval()
    c2w_host = sc["c2w"].pin_memory()
    c2w_dev = c2w_host.to(dev)
    frame_host = torch.empty(cam.h, cam.w, 3, dtype=torch.float32).pin_memory() if rank == 0 else None

    # N > 1: the bands are composited straight into rank 0's frame buffer over NVLink (parallel.SharedFrame);
    # `--band-gather allgather` selects the round-1 padded all_gather of the bands instead
    shared = P.SharedFrame(cam, dev) if (world > 1 and args.band_gather == "root") else None

    def frame(e2e):
        c2w = c2w_host.to(dev, non_blocking=True) if e2e else c2w_dev
        img = P.tile_sharded_render(r, c2w, cam, frame=shared)
        if e2e and rank == 0:  # what the viewer does with a frame (viser_viewer.py:119-139: `.cpu()`)
            frame_host.copy_(img, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        return img

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
        frame(True)
        frame(False)
    with ClockSampler(phys_gpu_index(local_rank)) as clk:
        l0 = capi.lib.gs3d_launch_count()
        ms = timed(lambda: frame(False), args.steps)
        launches = capi.lib.gs3d_launch_count() - l0
        ms_e2e = timed(lambda: frame(True), args.steps)
    # single-GPU image of the same frame for the sharding check (bit-exact: tiles are independent)
    full = frame(False)
    with torch.no_grad():
        ref_img = r(c2w_dev, cam)
    shard_err = float((full - ref_img).abs().max()) if full is not None else None  # (root only with SharedFrame)
    n_dub = r.total_dub_gaussians
    line = {
        "metric": "forward FPS (6M Gaussians SH3 @3840x2160, tile rows sharded)", "value": 1000.0 * args.steps / ms,
        "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_desc(name, r.N, C, cam, 1), "n_dub": n_dub, "views_per_step": 1,
                   "parallelism": f"tiles{world}" if world > 1 else "single"},
        "impl_details": {"sharding": (f"tile rows sharded over {world} rank(s); " +
                                      ("bands stored straight into rank 0's frame buffer over NVLink (symmetric "
                                       "memory, two barriers per frame)" if shared is not None else
                                       "bands gathered on every rank (padded all_gather)")),
                         "l2_policy": "inputs larger than L2 (parameters 1.4 GB, duplicates 1 GB vs 126 MB L2)",
                         "exact_decisions": not args.no_exact},
        "sharded_vs_single_gpu_max_abs": shard_err,
        "e2e": {"value": 1000.0 * args.steps / ms_e2e, "unit": "frames/s", "h2d_bytes_per_step": 48,
                "d2h_bytes_per_step": int(cam.h * cam.w * 12)},
        "gpu_launches": int(launches), "clocks": clk.summary(),
    }
    if args.check and world == 1:
        line["parity"] = parity_block(name, args, torch)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def parity_block(name, args, torch):
    """--check: whole-path parity of the benchmarked workload against the reference GPU flow (test
    infrastructure: oracle/fullsize_check.py + the real reference extension in oracle/_ref)."""
    try:
        from oracle import fullsize_check as F
        from oracle import ref_gpu

        ext = ref_gpu.load_reference_extension()
        if ext is None:
            return {"unavailable": "oracle/_ref/_gs_ref*.so absent"}
        res = F.compare_whole_path(ext, "cfg2" if name == "cfg4" else name, N=args.n_gaussians, seed=0,
                                   backward=name != "cfg5")
        res["pass"] = F.passes(res)
        return res
    except Exception as e:
        return {"error": str(e)}


FFMA_PEAK_TFLOPS = 69.2  # profiles/r1_microbench_pipes.txt: FFMA loop on one B200 at 1965 MHz (no FP32 peak in MEASURED_PEAKS)


def stage_rooflines(kernels_ms, N, C, cam, n_dub, views, staged, pairs, ncu_ok, touched_rows=0):
    """One roofline object per stage (DESIGN.md 'Kernels' states the per-unit figures).

    K1 / K2 / K4b are HBM-bound streaming kernels: achieved = algorithmic bytes / CUDA-event time against the
    measured copy bandwidth.  K3 / K4a are bound by FP32 issue + shared memory, not by HBM: their entry is
    bound = "issue", achieved = FLOP of the pairs actually evaluated / time against the measured FFMA rate; the
    HBM figure (bytes per STAGED duplicate) is kept beside it.  `views` launches per step are summed in
    kernels_ms, so per-launch quantities are multiplied by it."""
    hbm_peak, peak_src = peaks()
    CC = C * C
    px = cam.w * cam.h
    staged_fwd, staged_bwd = staged  # per step (all views of this rank)
    pairs_fwd, pairs_bwd = pairs
    per_dup = 4 + 48 + 12 * CC  # id + staging record + SH row
    algo = {
        # read mean 12 + qvec 16 + svec 12 + alpha 4; write the 48-byte record + rect 16 + depth 4 + mask 1
        "K1_project_cull": views * N * 89,  # (the kernel's own necessary traffic is 44 + 48 record + 21 = 113 B)
        # own algorithm (DESIGN 'K2'): per Gaussian 4 depth passes (histogram 4 B + scatter: 4 -> 8, 8 -> 8, 8 -> 8,
        # 8 + 16 rect gather -> 4 + 8) + count 8 + emit 12 = 116 B; per duplicate emit 8 + pass 0 (4 + 8 + 8) + last
        # pass (4 + 8 + 4, keys dropped, ranges extracted in the same kernel) = 44 B
        "K2_binning": views * N * 116 + n_dub * views * 44,
        "K3_composite_fwd": staged_fwd * per_dup + views * px * 12,
        "K4a_composite_bwd": staged_bwd * (per_dup + 4 * (7 + 3 * CC)) + views * px * 36,
        # every row: mask + the three upstream gradients (29 B); rows with a gradient (the marked ones) also read the
        # 44 B of parameters and read-modify-write 44 B of leaf gradients
        "K4b_project_bwd": views * N * (28 + 1) + touched_rows * (44 + 2 * 44),
    }
    # FLOP of the compositing kernels: 14 per tested pair; a contributing pair adds 6 C^2 + 12
    # (forward) / recompute + gradient arithmetic + the 6 C^2 SH outer product (backward)
    flops = {
        "K3_composite_fwd": 256 * staged_fwd * 14 + pairs_fwd * (6 * CC + 12),
        "K4a_composite_bwd": 256 * staged_bwd * 14 + pairs_bwd * (6 * CC + 12 + 6 * CC + 70),
    }
    ncu = {}
    tfile = ROOT / "profiles" / "ncu_traffic.json"
    if ncu_ok and tfile.exists():
        try:
            ncu = json.loads(tfile.read_text())
        except Exception:
            ncu = {}
    out = {}
    meta = ncu.get("_meta") or {}
    if meta.get("src_sha256_16"):  # was the committed ncu capture taken from the kernel sources that are built now?
        import hashlib

        h = hashlib.sha256()
        for f in sorted((ROOT / "gaussian_splatting_3d_b200" / "csrc").glob("*.cu*")) + [ROOT / "include" / "gs3d_b200.h"]:
            h.update(f.name.encode())
            h.update(f.read_bytes())
        now = h.hexdigest()[:16]
        out["_ncu_capture"] = {"src_sha256_16": meta["src_sha256_16"], "current_src_sha256_16": now,
                               "same_sources": now == meta["src_sha256_16"], "capture": meta.get("capture")}
    for k, ms in kernels_ms.items():
        if k not in algo or ms <= 0:
            continue
        t = ncu.get(k) or {}
        gbs = algo[k] / (ms * 1e-3) / 1e9
        if k in flops:
            tf = flops[k] / (ms * 1e-3) / 1e12
            o = {"kernel": k, "bound": "issue", "achieved": tf, "peak": FFMA_PEAK_TFLOPS, "unit": "TFLOP/s",
                 "frac": tf / FFMA_PEAK_TFLOPS, "peak_source": "measured FFMA loop (profiles/r1_microbench_pipes.txt)",
                 "flop": flops[k], "pairs_tested": 256 * (staged_fwd if "fwd" in k else staged_bwd),
                 "pairs_contributing": pairs_fwd if "fwd" in k else pairs_bwd,
                 "hbm": {"achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                         "algorithmic_bytes": algo[k]},
                 "staged_duplicates": staged_fwd if "fwd" in k else staged_bwd, "n_dub": n_dub}
        else:
            o = {"kernel": k, "bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s",
                 "frac": gbs / hbm_peak, "peak_source": peak_src, "algorithmic_bytes": algo[k]}
        o["launch_ms"] = ms / max(views, 1)
        o["traffic"] = t.get("traffic")
        o["traffic_source"] = t.get("source")
        if t.get("traffic") and algo[k]:
            o["traffic_over_algorithmic"] = t["traffic"] * views / algo[k]
        util = {m: t[m] for m in ("issue_slot_pct", "fma_pipe_pct", "lsu_pipe_pct", "smem_wavefront_pct",
                                  "warps_active_pct", "dram_pct") if t.get(m) is not None}
        if util:
            o["ncu_utilisation_pct_of_peak"] = util
        out[k] = o
    kk = {k: v for k, v in kernels_ms.items() if k in out and not k.startswith("_")}
    dom = max(kk, key=kk.get) if kk else None
    return out, dom


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
    cfg4, cfg5 = name == "cfg4", name == "cfg5"
    scene = "cfg2" if cfg4 else name
    cam = S.make_camera(scene)
    cpu = None
    if not args.no_cpu_baseline:
        try:
            cpu = cpu_baseline(args, scene)
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
    sc = S.make_scene(scene, seed=0, N=args.n_gaussians)
    C = sc["C"]
    ref = ref_gpu.ReferenceGPURenderer(ext, sc, dev, C)
    poses = S.ring_cameras(8) if cfg4 else [sc["c2w"]]  # cfg 4: the same 8 cameras per step, on one GPU
    c2w_hosts = [c.pin_memory() for c in poses]
    tgt_hosts = [S.make_target(cam, i).pin_memory() for i in range(len(poses))]
    c2w_devs, tgt_devs = [c.to(dev) for c in c2w_hosts], [t.to(dev) for t in tgt_hosts]
    c2w_host, tgt_host, c2w_dev = c2w_hosts[0], tgt_hosts[0], c2w_devs[0]

    def step(e2e):
        if cfg5:  # a render: forward only, the frame is read back in the end-to-end form
            with torch.no_grad():
                out = ref.forward(c2w_host.to(dev, non_blocking=True) if e2e else c2w_dev, cam)
            return out.cpu() if e2e else out
        ref.zero_grad()
        total = None
        for v in range(len(poses)):
            c2w = c2w_hosts[v].to(dev, non_blocking=True) if e2e else c2w_devs[v]
            tgt = tgt_hosts[v].to(dev, non_blocking=True) if e2e else tgt_devs[v]
            out = ref.forward(c2w, cam)
            loss = ((out - tgt) ** 2).mean()
            loss.backward()  # autograd accumulates over the step's views
            total = loss.detach() if total is None else total + loss.detach()
        return float(total.item()) if e2e else total

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
    nv = len(poses)
    if cfg5:
        base["metric"], base["unit"] = "forward FPS (6M Gaussians SH3 @3840x2160, tile rows sharded)", "frames/s"
    base["scaling"] = "strong" if (cfg4 or cfg5) else "weak"
    base.update({
        "value": nv * 1000.0 * args.steps / ms, "ms_per_step": ms / args.steps,
        "fwd_fps": 1000.0 * args.steps / ms_fwd,
        "config": {"workload": workload_desc(name, ref.params['mean'].shape[0], C, cam, nv),
                   "n_dub": ref.total_dub_gaussians, "views_per_step": nv, "parallelism": "single"},
        "impl_details": {"what": "reference CUDA extension (sm_100a build of /root/reference/gs/src, -DNDEBUG; its "
                                 "device printf output discarded) behind the reference's torch-level flow, 1 GPU"},
        "e2e": {"value": nv * 1000.0 * args.steps / ms_e2e, "unit": base["unit"],
                "h2d_bytes_per_step": 48 if cfg5 else int(nv * (c2w_host.numel() * 4 + tgt_host.numel() * 4)),
                "d2h_bytes_per_step": int(cam.h * cam.w * 12) if cfg5 else 4},
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
