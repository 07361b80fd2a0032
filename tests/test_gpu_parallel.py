"""GPU tests of the sharded paths.  Single-GPU: the union of independently rendered tile-row bands
equals the full render bit for bit (tiles are independent).  Two GPUs (skipped on a 1-GPU box):
tile-sharded render == single-GPU render, and view-sharded training == serial sum of gradients."""
import os

import pytest
import torch

from conftest import free_port
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussian_splatting_3d_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _renderer(dev, name="cfg3", n=60_000, C=3, seed=0):
    sc = S.make_scene(name, seed=seed, N=n, C=C)
    r = S.renderer_from_scene(sc, S.make_cfg(device=dev, sh_order=C))
    return r, sc, S.make_camera(name)


def test_bands_compose_to_full_frame_single_gpu():
    from gaussian_splatting_3d_b200 import parallel as P

    dev = "cuda:0"
    r, sc, cam = _renderer(dev)
    c2w = sc["c2w"].to(dev)
    with torch.no_grad():
        full = r(c2w, cam).clone()
    img_all, k1 = P.render_band(r, c2w, cam)
    assert torch.equal(img_all, full)
    nth = (cam.h + 15) // 16
    rc = P.row_duplicate_counts(k1["tl"], k1["br"], nth)
    assert int(rc.sum()) == r.total_dub_gaussians
    for world in (2, 8):
        bands = P.balanced_bands(rc, world)
        acc = torch.zeros_like(full)
        for a, b in bands:
            img, _ = P.render_band(r, c2w, cam, a, b, k1=k1)
            y0, y1 = a * 16, min(b * 16, cam.h)
            acc[y0:y1] = img[y0:y1]  # (rows outside the band are undefined: the buffer is not filled)
        assert torch.equal(acc, full)
        loads = [int(rc[a:b].sum()) for a, b in bands]
        assert max(loads) <= sum(loads) / world + int(rc.max())


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    from gaussian_splatting_3d_b200 import parallel as P

    r, sc, cam = _renderer(dev)
    r.train()
    c2w = sc["c2w"].to(dev)
    full = P.tile_sharded_render(r, c2w, cam)
    with torch.no_grad():
        single = r(c2w, cam)
    ok_render = bool(torch.equal(full, single))
    # same frame with the bands stored straight into rank 0's frame buffer (symmetric memory, no gather step);
    # twice: the second frame must wait until the root is done with the first
    shared = P.SharedFrame(cam, dev)
    for _ in range(2):
        img = P.tile_sharded_render(r, c2w, cam, frame=shared)
        if rank == 0:
            ok_render = ok_render and bool(torch.equal(img, single))
        else:
            ok_render = ok_render and img is None
    # view-sharded step over 4 cameras
    import math
    c2ws = []
    for i in range(4):
        a = math.radians(3.0 * (i - 1.5))
        c2ws.append(torch.tensor([[math.cos(a), 0, math.sin(a), 0], [0, 1, 0, 0], [-math.sin(a), 0, math.cos(a), 0]],
                                 dtype=torch.float32, device=dev))
    targets = [S.make_target(cam, i).to(dev) for i in range(4)]
    r._reset_adc_buffers()  # the renders above already counted one view
    flat = P.FlatGradients(r)
    P.view_sharded_step(r, flat, c2ws, cam, targets)
    sharded = flat.flat.clone()
    cnt_sharded = r.cnt.clone()
    # serial reference on this rank: all four views
    r2, _, _ = _renderer(dev)
    r2.train()
    flat2 = P.FlatGradients(r2)
    flat2.zero()
    for i in range(4):
        ((r2(c2ws[i], cam) - targets[i]) ** 2).mean().backward()
    rel = float((sharded - flat2.flat).norm() / flat2.flat.norm())
    q.put((rank, ok_render, rel, bool(torch.equal(cnt_sharded, r2.cnt))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_tile_sharded_render_and_view_sharded_training():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_render, rel, cnt_ok in res:
        assert ok_render, f"rank {rank}: tile-sharded frame differs from the single-GPU frame"
        assert rel < 1e-4, f"rank {rank}: view-sharded gradient sum differs (rel {rel:.2e})"
        assert cnt_ok


def _worker_fused(rank, world, port, q, use_multicast):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    import math

    from gaussian_splatting_3d_b200 import parallel as P

    r, sc, cam = _renderer(dev, name="cfg2", n=80_000, C=4)
    r.train()
    c2ws = []
    for i in range(4):
        a = math.radians(3.0 * (i - 1.5))
        c2ws.append(torch.tensor([[math.cos(a), 0, math.sin(a), 0], [0, 1, 0, 0], [-math.sin(a), 0, math.cos(a), 0]],
                                 dtype=torch.float32, device=dev))
    targets = [S.make_target(cam, i).to(dev) for i in range(4)]
    flat = P.FlatGradients(r, fused=True, use_multicast=use_multicast).attach(r)
    info = (flat.fused, flat.multicast_ptr is not None, len(flat.peer_ptrs or []))
    for _ in range(2):  # twice: the buffers must be re-zeroed and re-synchronised correctly
        P.view_sharded_step(r, flat, c2ws, cam, targets)
    fused_sum = flat.flat.clone()
    # serial reference on this rank: all four views with plain autograd accumulation
    r2, _, _ = _renderer(dev, name="cfg2", n=80_000, C=4)
    r2.train()
    flat2 = P.FlatGradients(r2)
    flat2.zero()
    for i in range(4):
        ((r2(c2ws[i], cam) - targets[i]) ** 2).mean().backward()
    n_sh = flat.n_sh
    rel_sh = float((fused_sum[:n_sh] - flat2.flat[:n_sh]).norm() / flat2.flat[:n_sh].norm())
    rel_rest = float((fused_sum[n_sh:] - flat2.flat[n_sh:]).norm() / flat2.flat[n_sh:].norm())
    q.put((rank, info, rel_sh, rel_rest))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("use_multicast", [False, True])
def test_two_gpu_fused_gradient_exchange(use_multicast):
    """SH gradients reduced into both ranks' buffers by the backward kernel itself (peer stores or
    NVSwitch multimem) + small all-reduce == serial sum over the four views."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker_fused, args=(r, 2, port, q, use_multicast)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, info, rel_sh, rel_rest in res:
        print(f"[fused dp rank {rank}] fused={info[0]} multicast={info[1]} peers={info[2]} "
              f"rel_sh={rel_sh:.2e} rel_rest={rel_rest:.2e}")
        assert info[0] and info[2] == 2
        assert rel_sh < 1e-4 and rel_rest < 1e-4


def _worker_sparse(rank, world, port, q, mode):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    import math

    from gaussian_splatting_3d_b200 import parallel as P

    r, sc, cam = _renderer(dev, name="cfg2", n=80_000, C=4)
    r.train()
    c2ws = []
    for i in range(4):
        a = math.radians(3.0 * (i - 1.5))
        c2ws.append(torch.tensor([[math.cos(a), 0, math.sin(a), 0], [0, 1, 0, 0], [-math.sin(a), 0, math.cos(a), 0]],
                                 dtype=torch.float32, device=dev))
    targets = [S.make_target(cam, i).to(dev) for i in range(4)]
    flat = P.FlatGradients(r, sparse=(mode == "sparse"), push=(mode.startswith("push")),
                           pull=(mode.startswith("pull")), use_multicast=not mode.endswith("-peer")).attach(r)
    assert flat.sparse == (mode == "sparse") and flat.push == (mode != "sparse") and flat.pull == mode.startswith("pull")
    for _ in range(3):  # repeatedly: marks and buffers must be reset correctly
        P.view_sharded_step(r, flat, c2ws, cam, targets)
    got = flat.flat.clone()
    r2, _, _ = _renderer(dev, name="cfg2", n=80_000, C=4)
    r2.train()
    flat2 = P.FlatGradients(r2)
    flat2.zero()
    for i in range(4):
        ((r2(c2ws[i], cam) - targets[i]) ** 2).mean().backward()
    want = flat2.flat
    rel = float((got - want).norm() / want.norm())
    # rows outside the union must be exactly zero in the serial sum too (nothing was dropped)
    N = r.mean.size(0)
    nz_rows = int((want[: flat.n_sh].view(N, -1).abs().sum(1) > 0).sum())
    union = flat.last_union_rows if mode == "sparse" else int(flat.union.count_nonzero())
    q.put((rank, rel, union, nz_rows, N))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode", ["sparse", "push", "push-peer", "pull", "pull-peer"])
def test_two_gpu_sparse_gradient_exchange(mode):
    """Exchange of the touched rows only -- 'sparse': union marks + NCCL all-reduce of the packed rows
    (gs3d_rows_gather / gs3d_rows_scatter); 'push': every rank adds its rows into every rank's result
    buffer over the NVSwitch multicast address ('push-peer': peer by peer); 'pull': union marks, then
    multimem.ld_reduce + multimem.st per union row ('pull-peer': peer loads / stores) -- == serial sum over the
    four views; the union covers every non-zero row."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker_sparse, args=(r, 2, port, q, mode)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, rel, union_rows, nz_rows, N in res:
        print(f"[{mode} dp rank {rank}] union rows {union_rows} of {N} (non-zero SH rows in the serial sum: {nz_rows}) "
              f"rel={rel:.2e}")
        assert rel < 1e-4
        assert nz_rows <= union_rows < N


def _worker_adc(rank, world, port, q):
    """view_sharded_step -> adaptive_control(force=True) (split + prune + alpha reset: N changes, the ADC
    buffers are replaced) -> view_sharded_step again, on both ranks."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = f"cuda:{rank}"
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(dev))
    import math

    from gaussian_splatting_3d_b200 import parallel as P

    sc = S.make_scene("cfg3", seed=0, N=40_000, C=2)
    cam = S.make_camera("cfg3")
    r = S.renderer_from_scene(sc, S.make_cfg(device=dev, sh_order=2, warm_up=0, pos_grad_thresh=1e-7))
    r.train()
    r.fuse_adc = True
    c2ws = []
    for i in range(4):
        a = math.radians(3.0 * (i - 1.5))
        c2ws.append(torch.tensor([[math.cos(a), 0, math.sin(a), 0], [0, 1, 0, 0], [-math.sin(a), 0, math.cos(a), 0]],
                                 dtype=torch.float32, device=dev))
    targets = [S.make_target(cam, i).to(dev) for i in range(4)]
    flat = P.FlatGradients(r, pull=True).attach(r)
    P.view_sharded_step(r, flat, c2ws, cam, targets)
    cnt1, gm1 = r.cnt.clone(), r.grad_mean.clone()
    n0 = r.N
    torch.manual_seed(123)  # the split draws torch.randn: same stream on both ranks
    r.adaptive_control(0, force=True)
    n1 = r.N
    stale = None
    try:
        flat.zero()
    except RuntimeError as e:
        stale = str(e)
    flat = P.FlatGradients(r, pull=True).attach(r)  # the owner rebuilds the buffers for the new parameters
    P.view_sharded_step(r, flat, c2ws, cam, targets)
    # serial reference of the second step's statistics on this rank
    r2 = S.renderer_from_scene({k: (getattr(r, k).data.cpu() if k in ("mean", "qvec", "svec_before_activation", "sh_coeffs",
                                                                    "alpha_before_activation") else v)
                                for k, v in sc.items()},
                               S.make_cfg(device=dev, sh_order=2, warm_up=0))
    r2.now_C = r.now_C
    r2.train()
    r2.fuse_adc = True
    for i in range(4):
        ((r2(c2ws[i], cam) - targets[i]) ** 2).mean().backward()
    gm_rel = float((r.grad_mean.double() - r2.grad_mean.double()).norm() / r2.grad_mean.double().norm().clamp_min(1e-30))
    q.put((rank, n0, n1, stale is not None, bool(torch.equal(r.cnt, r2.cnt)), gm_rel, int(cnt1.max()), float(gm1.sum()),
           int(r.mean.data.sum().item() * 1000)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_view_sharded_training_across_adaptive_control():
    """sync_adc's per-renderer baseline and the attached gradient views must not survive a change
    of N -- statistics after the density-control step equal the serial ones, stale buffers raise."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker_adc, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, n0, n1, stale_raised, cnt_ok, gm_rel, cnt_max, gm_sum, chk in res:
        print(f"[adc dp rank {rank}] N {n0} -> {n1}, stale buffers rejected: {stale_raised}, cnt ok {cnt_ok}, "
              f"grad_mean rel {gm_rel:.2e}")
        assert n1 != n0 and stale_raised
        assert cnt_max == 4 and gm_sum > 0      # first step: four views counted on every rank
        assert cnt_ok and gm_rel < 1e-3
    assert res[0][1:3] == res[1][1:3] and res[0][-1] == res[1][-1]  # both ranks hold the same model


@pytest.mark.gpu
def test_rows_gather_scatter_roundtrip():
    from gaussian_splatting_3d_b200 import ops

    dev = "cuda:0"
    g = torch.Generator().manual_seed(3)
    N = 1000
    blocks = [torch.randn(N, 3, 16, generator=g).to(dev), torch.randn(N, 3, generator=g).to(dev),
              torch.randn(N, 4, generator=g).to(dev), torch.randn(N, generator=g).to(dev)]
    idx = torch.randperm(N, generator=g)[:137].sort().values.int().to(dev)
    packed = torch.full((137, 60), float("nan"), device=dev)
    ops.rows_gather(blocks, idx, packed)
    want = torch.cat([b.view(N, -1)[idx.long()] for b in blocks], 1)
    assert torch.equal(packed[:, :56], want) and bool((packed[:, 56:] == 0).all())
    outs = [torch.zeros_like(b) for b in blocks]
    ops.rows_scatter(outs, idx, packed * 2)
    for o, b in zip(outs, blocks):
        ref = torch.zeros_like(b)
        ref[idx.long()] = 2 * b[idx.long()]
        assert torch.equal(o, ref)
    # empty index list is a no-op
    ops.rows_gather(blocks, idx[:0], packed[:0])


def test_band_kernels_match_torch_restatement():
    """gs3d_row_duplicate_counts / gs3d_clip_rects_to_rows against the integer torch restatement
    (parallel.row_duplicate_counts on CPU tensors, parallel.clip_rects_to_band)."""
    from gaussian_splatting_3d_b200 import ops
    from gaussian_splatting_3d_b200 import parallel as P

    g = torch.Generator().manual_seed(11)
    N, nth, ntw = 20_000, 47, 63
    tl = torch.stack([torch.randint(0, ntw, (N,), generator=g), torch.randint(0, nth, (N,), generator=g)], 1).int()
    ext = torch.stack([torch.randint(-1, 6, (N,), generator=g), torch.randint(-1, 9, (N,), generator=g)], 1).int()
    br = tl + ext                                      # some rects are empty (extent -1)
    br[:, 0].clamp_(max=ntw - 1)
    br[:, 1].clamp_(max=nth - 1)
    depth = torch.rand(N, 1, generator=g)
    want_rows = P.row_duplicate_counts(tl, br, nth)     # CPU restatement
    got_rows = ops.row_duplicate_counts(tl.cuda(), br.cuda(), nth)
    assert torch.equal(got_rows.cpu(), want_rows)
    for r0, r1 in ((0, nth), (5, 17), (30, 31), (46, 47), (10, 10)):
        ctl, cbr, n_want = P.clip_rects_to_band(tl, br, r0, r1)
        keep = ((cbr[:, 0] - ctl[:, 0] + 1) > 0) & ((cbr[:, 1] - ctl[:, 1] + 1) > 0)
        gtl, gbr, gd, gidx, n_got = ops.clip_rects_to_rows(tl.cuda(), br.cuda(), depth.cuda(), r0, r1)
        assert n_got == int(n_want)
        assert torch.equal(gidx.cpu().long(), torch.nonzero(keep).view(-1))
        assert torch.equal(gtl.cpu(), ctl[keep]) and torch.equal(gbr.cpu(), cbr[keep])
        assert torch.equal(gd.cpu(), depth[keep])
