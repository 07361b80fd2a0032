"""CPU tests of the multi-GPU host logic: band partitioning for tile-sharded rendering and the
flat-gradient all-reduce of view-sharded training over a world_size-2 gloo group."""
import os

import torch

from conftest import free_port
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussian_splatting_3d_b200 import parallel as P


def test_row_counts_and_bands_cover_everything():
    g = torch.Generator().manual_seed(0)
    n, nth, ntw = 5000, 53, 82
    tlx = torch.randint(0, ntw, (n,), generator=g)
    tly = torch.randint(0, nth, (n,), generator=g)
    brx = torch.minimum(tlx + torch.randint(0, 5, (n,), generator=g), torch.tensor(ntw - 1))
    bry = torch.minimum(tly + torch.randint(0, 5, (n,), generator=g), torch.tensor(nth - 1))
    tl = torch.stack([tlx, tly], 1).int()
    br = torch.stack([brx, bry], 1).int()
    tl[::11] = torch.tensor([0, 0], dtype=torch.int32)  # culled-style empty rects
    br[::11] = torch.tensor([-1, -1], dtype=torch.int32)
    rc = P.row_duplicate_counts(tl, br, nth)
    # brute force
    want = torch.zeros(nth, dtype=torch.long)
    for i in range(n):
        w = int(br[i, 0] - tl[i, 0] + 1)
        for y in range(int(tl[i, 1]), int(br[i, 1]) + 1):
            want[y] += w
    assert torch.equal(rc, want)
    total = int(((br[:, 0] - tl[:, 0] + 1).clamp(min=0).long() * (br[:, 1] - tl[:, 1] + 1).clamp(min=0).long()).sum())
    assert int(rc.sum()) == total
    for world in (1, 2, 3, 4, 8):
        bands = P.balanced_bands(rc, world)
        assert len(bands) == world and bands[0][0] == 0 and bands[-1][1] == nth
        assert all(bands[i][1] == bands[i + 1][0] for i in range(world - 1))
        parts = []
        for a, b in bands:
            tlc, brc, nb = P.clip_rects_to_band(tl, br, a, b)
            parts.append(int(nb))
            assert int(nb) == int(rc[a:b].sum())
        assert sum(parts) == total
        if world > 1:
            assert max(parts) <= total / world + int(rc.max())  # balanced to within one row


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class M(torch.nn.Module):
        pass

    m = M()
    g = torch.Generator().manual_seed(7)
    n = 50
    m.mean = torch.nn.Parameter(torch.randn(n, 3, generator=g))
    m.qvec = torch.nn.Parameter(torch.randn(n, 4, generator=g))
    m.svec_before_activation = torch.nn.Parameter(torch.randn(n, 3, generator=g))
    m.sh_coeffs = torch.nn.Parameter(torch.randn(n, 3, 4, generator=g))
    m.alpha_before_activation = torch.nn.Parameter(torch.randn(n, generator=g))
    m.split_reduction = "mean"
    m.grad_mean = torch.zeros(n)
    m.cnt = torch.zeros(n, dtype=torch.int32)
    flat = P.FlatGradients(m)
    flat.zero()
    views = P.shard_views(8, rank, world)
    # a stand-in "renderer": loss depends on the view index so ranks contribute different gradients
    for v in views:
        loss = sum(((p * (v + 1)) ** 2).sum() for p in flat.params)
        loss.backward()
        m.grad_mean += float(v + 1)
        m.cnt += 1
    flat.all_reduce()
    P.sync_adc(m)
    # (numpy arrays travel through the queue BY VALUE; torch tensors travel as shared-memory handles served by this
    # process, and the parent's get() raced with this worker's exit: one run in eight failed with FileNotFoundError)
    q.put((rank, views, flat.flat.clone().numpy(), m.grad_mean.clone().numpy(), m.cnt.clone().numpy(),
           [p.grad.data_ptr() == vw.data_ptr() for p, vw in zip(flat.params, flat.views)]))
    dist.barrier()
    dist.destroy_process_group()


def test_view_sharded_gradient_allreduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = [tuple(torch.from_numpy(x) if hasattr(x, "dtype") and not torch.is_tensor(x) else x for x in t) for t in res]
    (r0, v0, f0, gm0, c0, al0), (r1, v1, f1, gm1, c1, al1) = res
    assert v0 == [0, 2, 4, 6] and v1 == [1, 3, 5, 7]
    assert all(al0) and all(al1)  # .grad aliases the flat buffer
    assert torch.equal(f0, f1)  # every rank holds the same summed gradient
    # single-process reference: sum over all 8 views
    g = torch.Generator().manual_seed(7)
    n = 50
    ps = [torch.randn(n, 3, generator=g), torch.randn(n, 4, generator=g), torch.randn(n, 3, generator=g),
          torch.randn(n, 3, 4, generator=g), torch.randn(n, generator=g)]
    scale = sum(2.0 * (v + 1) ** 2 for v in range(8))
    # flat layout: [sh_coeffs | mean | qvec | svec | alpha], every block padded to a multiple of 4 floats (16 bytes)
    def padded(t):
        v = (t * scale).reshape(-1)
        return torch.cat([v, torch.zeros((-v.numel()) % 4)])

    want = torch.cat([padded(p) for p in (ps[3], ps[0], ps[1], ps[2], ps[4])])
    assert torch.allclose(f0, want, rtol=1e-5, atol=1e-5)
    assert torch.equal(gm0, gm1) and float(gm0[0]) == float(sum(range(1, 9)))
    assert torch.equal(c0, c1) and int(c0[0]) == 8


def _worker_adc_baseline(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class M:
        pass

    m = M()
    m.split_reduction, m.split_type = "mean", "2d_mean_grad"
    m.grad_mean, m.cnt, m._adc_epoch = torch.zeros(6), torch.zeros(6, dtype=torch.int32), 1
    out = []
    # step 1: rank-distinct contributions
    m.grad_mean += float(rank + 1)
    m.cnt += 1
    P.sync_adc(m)
    out.append((m.grad_mean.clone(), m.cnt.clone()))
    # step 2 accumulates on top of the synchronised state
    m.grad_mean += float(rank + 1)
    m.cnt += 1
    P.sync_adc(m)
    out.append((m.grad_mean.clone(), m.cnt.clone()))
    # density control: N changes and the buffers are replaced (SHRenderer._reset_adc_buffers bumps the epoch)
    m.grad_mean, m.cnt, m._adc_epoch = torch.zeros(9), torch.zeros(9, dtype=torch.int32), 2
    m.grad_mean += float(rank + 1)
    m.cnt += 1
    P.sync_adc(m)
    out.append((m.grad_mean.clone(), m.cnt.clone()))
    # an alpha reset keeps N but re-zeroes the buffers: the stale baseline must not be subtracted
    m.grad_mean, m.cnt, m._adc_epoch = torch.zeros(9), torch.zeros(9, dtype=torch.int32), 3
    m.grad_mean += float(rank + 1)
    m.cnt += 1
    P.sync_adc(m)
    out.append((m.grad_mean.clone(), m.cnt.clone()))
    # split_type mean_grad: the statistic comes from the already exchanged gradient -> identical on all ranks, not summed
    m.split_type = "mean_grad"
    m.grad_mean += 5.0
    m.cnt += 1
    P.sync_adc(m)
    out.append((m.grad_mean.clone(), m.cnt.clone()))
    q.put((rank, [(gm.numpy(), c.numpy()) for gm, c in out]))  # by value (see _worker)
    dist.barrier()
    dist.destroy_process_group()


def test_sync_adc_baseline_follows_buffer_resets_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=_worker_adc_baseline, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = [(rk, [(torch.from_numpy(gm), torch.from_numpy(c)) for gm, c in out]) for rk, out in res]
    for (gm0, c0), (gm1, c1) in zip(res[0][1], res[1][1]):
        assert torch.equal(gm0, gm1) and torch.equal(c0, c1)
    o = res[0][1]
    assert float(o[0][0][0]) == 3.0 and int(o[0][1][0]) == 2      # 1 + 2, two views
    assert float(o[1][0][0]) == 6.0 and int(o[1][1][0]) == 4
    assert o[2][0].numel() == 9 and float(o[2][0][0]) == 3.0 and int(o[2][1][0]) == 2   # fresh after the split
    assert float(o[3][0][0]) == 3.0 and int(o[3][1][0]) == 2                          # fresh after the alpha reset
    assert float(o[4][0][0]) == 8.0 and int(o[4][1][0]) == 4                          # +5 once, cnt +1 per rank


def test_stale_gradient_views_are_rejected():
    """a renderer that replaced its parameters drops the attached gradient views, and the
    FlatGradients that owned them refuses to be used again."""

    class M(torch.nn.Module):
        pass

    m = M()
    for n, shape in (("mean", (5, 3)), ("qvec", (5, 4)), ("svec_before_activation", (5, 3)), ("sh_coeffs", (5, 3, 4)),
                     ("alpha_before_activation", (5,))):
        setattr(m, n, torch.nn.Parameter(torch.zeros(shape)))
    m.grad_buffers = None
    flat = P.FlatGradients(m).attach(m)
    flat.zero()
    m.mean = torch.nn.Parameter(torch.zeros(7, 3))  # what _set_params does ...
    m.grad_buffers = None                           # ... including dropping the views
    try:
        flat.zero()
        raised = False
    except RuntimeError:
        raised = True
    assert raised
