"""Multi-GPU paths (new capability: the reference has no distributed code).
One process per GPU, torch.distributed (NCCL on GPUs, gloo in the CPU tests) for the plumbing.

(i)  View-sharded training (cfg 4): parameters replicated, each rank renders its share of the
     step's cameras forward+backward, gradients are summed with ONE all-reduce over a flat buffer
     that the parameters' `.grad`s alias (236 B/Gaussian at C = 4); the ADC statistics follow
     (grad_mean: sum or max; cnt: sum).
(ii) Tile-sharded rendering (cfg 5): projection is replicated (cheap), each rank bins and
     composites one band of tile rows -- bands balanced by per-row duplicate counts -- and the
     bands are gathered into the full frame.  No reduction is needed: tiles are independent.
"""
import torch
import torch.distributed as dist

_PARAMS = ("mean", "qvec", "svec_before_activation", "sh_coeffs", "alpha_before_activation")


class FlatGradients:
    """One flat FP32 buffer holding every leaf gradient; the parameters' `.grad`s are views of it,
    so the data-parallel exchange needs no packing copy.  Layout: [sh_coeffs | mean | qvec | svec |
    alpha] -- the SH block (576 of 708 MB at C = 4) first, the small dense block after it.

    attach(renderer) makes the renderer's backward kernels write / accumulate straight into those
    views (`renderer.grad_buffers`); such steps must use `backward_into()` (torch.autograd.grad)
    instead of `loss.backward()`, because autograd's own accumulation would add the buffer to itself.

    fused=True (needs an initialised NCCL group): the buffer is allocated as torch symmetric memory,
    every rank maps every other rank's buffer (and the NVSwitch multicast address when available),
    and the compositing-backward kernel reduces its SH gradient rows into ALL ranks' buffers while it
    runs (gs3d_composite_sh_backward_peers).  exchange() then only all-reduces the small dense block.
    """

    ORDER = ("sh_coeffs", "mean", "qvec", "svec_before_activation", "alpha_before_activation")

    def __init__(self, module, fused=False, group=None, use_multicast=True):
        self.names = list(self.ORDER)
        self.params = [getattr(module, n) for n in self.names]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.group = group
        self.fused = bool(fused) and dist.is_initialized() and dist.get_world_size(group) > 1
        self.handle = None
        if self.fused:
            import torch.distributed._symmetric_memory as symm_mem

            self.flat = symm_mem.empty(total, dtype=torch.float32, device=dev)
            self.handle = symm_mem.rendezvous(self.flat, group if group is not None else dist.group.WORLD)
        else:
            self.flat = torch.empty(total, dtype=torch.float32, device=dev)
        self.flat.zero_()
        self.views = []
        off = 0
        for p in self.params:
            v = self.flat[off:off + p.numel()].view_as(p)
            p.grad = v
            self.views.append(v)
            off += p.numel()
        self.n_sh = self.params[0].numel()
        self.peer_ptrs = None
        self.multicast_ptr = None
        if self.fused:
            # the SH block starts at offset 0 of the symmetric buffer on every rank
            self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
            mc = int(self.handle.multicast_ptr) if use_multicast else 0
            self.multicast_ptr = mc if mc else None
        self.module = None

    def attach(self, renderer):
        bufs = dict(zip(self.names, self.views))
        if self.fused:
            bufs["sh_peer_ptrs"] = self.peer_ptrs
            bufs["sh_multicast_ptr"] = self.multicast_ptr
        renderer.grad_buffers = bufs
        self.module = renderer
        return self

    def detach(self):
        if self.module is not None:
            self.module.grad_buffers = None
            self.module = None

    def _barrier(self):
        if self.handle is not None:
            self.handle.barrier(channel=0)
        elif dist.is_initialized():
            dist.barrier(group=self.group)

    def zero(self):
        self.flat.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v  # optimisers / zero_grad(set_to_none) may have dropped the alias
        if self.fused:
            self._barrier()  # nobody may add into a buffer that is not zeroed yet

    def backward_into(self, loss):
        """Run backward for `loss`; the kernels add the leaf gradients into the flat buffer."""
        if self.module is None:
            loss.backward()  # plain autograd accumulation into the aliased .grad views
        else:
            torch.autograd.grad(loss, self.params)

    def exchange(self, average=False):
        """Make every rank's buffer hold the sum over ranks."""
        if not (dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return self.flat
        if self.fused and self.module is not None:
            self._barrier()  # every rank's in-kernel reductions into this buffer have landed
            dist.all_reduce(self.flat[self.n_sh:], op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        if average:
            self.flat.div_(dist.get_world_size(self.group))
        return self.flat

    def all_reduce(self, group=None, average=False):
        return self.exchange(average=average)


def shard_views(n_views, rank, world):
    """Indices of the step's cameras rendered by `rank` (round-robin; 8 cameras over 1/2/4/8 ranks)."""
    return list(range(rank, n_views, world))


def view_sharded_step(renderer, flat, c2ws, camera_info, targets, loss_fn=None, group=None):
    """One data-parallel training step over `c2ws` (the whole step's cameras, same list on every
    rank).  Returns this rank's summed loss tensor.  After the call every rank holds the SUM of the
    gradients over all cameras in the parameters' .grad (and in flat.flat)."""
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if loss_fn is None:
        loss_fn = lambda out, tgt: ((out - tgt) ** 2).mean()  # noqa: E731
    flat.zero()
    total = None
    for i in shard_views(len(c2ws), rank, world):
        out = renderer(c2ws[i], camera_info)
        loss = loss_fn(out, targets[i])
        flat.backward_into(loss)  # accumulates into the flat buffer
        total = loss.detach() if total is None else total + loss.detach()
    flat.exchange()
    sync_adc(renderer, group)
    return total


def sync_adc(renderer, group=None):
    """Make the ADC statistics identical on all ranks (sh_renderer.py:602-623 semantics over the
    union of the step's views): cnt is summed; grad_mean is summed ('mean') or max-reduced ('max').
    Each rank contributes only what it accumulated since the last sync."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    if not hasattr(renderer, "_adc_synced"):
        renderer._adc_synced = (torch.zeros_like(renderer.grad_mean), torch.zeros_like(renderer.cnt))
    gm0, cnt0 = renderer._adc_synced
    d_cnt = renderer.cnt - cnt0
    dist.all_reduce(d_cnt, op=dist.ReduceOp.SUM, group=group)
    renderer.cnt = cnt0 + d_cnt
    if renderer.split_reduction == "max":
        gm = renderer.grad_mean.clone()
        dist.all_reduce(gm, op=dist.ReduceOp.MAX, group=group)
    else:
        d_gm = renderer.grad_mean - gm0
        dist.all_reduce(d_gm, op=dist.ReduceOp.SUM, group=group)
        gm = gm0 + d_gm
    renderer.grad_mean = gm
    renderer._adc_synced = (gm.clone(), renderer.cnt.clone())


# ---------------------------------------------------------------- tile-sharded rendering


def row_duplicate_counts(aabb_topleft, aabb_bottomright, n_tiles_h):
    """Duplicates per tile row from the per-Gaussian rects (int64 [n_tiles_h]); difference array +
    prefix sum, integer arithmetic only, so every rank computes the same numbers."""
    tl, br = aabb_topleft.long(), aabb_bottomright.long()
    w = (br[:, 0] - tl[:, 0] + 1).clamp_(min=0)
    valid = (br[:, 1] >= tl[:, 1]) & (w > 0)
    w = w * valid
    diff = torch.zeros(n_tiles_h + 1, dtype=torch.long, device=tl.device)
    diff.index_add_(0, tl[:, 1].clamp(0, n_tiles_h), w)
    diff.index_add_(0, (br[:, 1] + 1).clamp(0, n_tiles_h), -w)
    return torch.cumsum(diff, 0)[:n_tiles_h]


def balanced_bands(row_counts, world):
    """Split tile rows into `world` contiguous bands with near-equal duplicate totals.
    -> list of (row_begin, row_end) covering [0, n_rows) in order; empty bands allowed."""
    rc = row_counts.tolist()
    n = len(rc)
    total = sum(rc)
    bounds = [0]
    acc, r = 0, 0
    for k in range(1, world):
        target = total * k / world
        while r < n and acc + rc[r] / 2 <= target:
            acc += rc[r]
            r += 1
        bounds.append(max(r, bounds[-1]))
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(world)]


def clip_rects_to_band(aabb_topleft, aabb_bottomright, row_begin, row_end):
    """Restrict rects to tile rows [row_begin, row_end); rects outside become empty (br < tl)."""
    tl = aabb_topleft.clone()
    br = aabb_bottomright.clone()
    tl[:, 1].clamp_(min=row_begin)
    br[:, 1].clamp_(max=row_end - 1)
    n = ((br[:, 0] - tl[:, 0] + 1).clamp(min=0).long() * (br[:, 1] - tl[:, 1] + 1).clamp(min=0).long()).sum()
    return tl, br, n


@torch.no_grad()
def render_band(renderer, c2w, camera_info, row_begin=None, row_end=None, k1=None):
    """Forward-render tile rows [row_begin, row_end) of one frame (None = all rows).  Returns the
    full-size [H,W,3] buffer in which only the band's rows are written, plus the K1 outputs."""
    from . import ops

    dev = renderer.mean.device
    cam, tile, C = camera_info, renderer.tile_size, renderer.now_C
    c2w = c2w.contiguous().float()
    if k1 is None:
        k1 = ops.project_cull_fused(
            renderer.mean.data, renderer.qvec.data, renderer.svec_before_activation.data,
            renderer.alpha_before_activation.data, renderer._svec_code, renderer._alpha_code, c2w, cam,
            renderer.frustum_culling_radius, renderer.skip_frustum_culling, renderer.tile_culling_radius, tile,
            cnt=None, want_records=True, want_activated=False)
    H, W = cam.h, cam.w
    nth = H // tile + (H % tile > 0)
    ntw = W // tile + (W % tile > 0)
    if row_begin is None:
        row_begin, row_end = 0, nth
    tl, br, n_band = clip_rects_to_band(k1["tl"], k1["br"], row_begin, row_end)
    n_band = int(n_band.item())
    ids = torch.empty(n_band, dtype=torch.int32, device=dev)
    start = torch.empty(nth * ntw, dtype=torch.int32, device=dev)
    end = torch.empty(nth * ntw, dtype=torch.int32, device=dev)
    ops.tile_culling_aabb_start_end(tl, br, ids, start, end, k1["depth"], nth, ntw, check_count=False)
    out = torch.zeros(H * W * 3, dtype=torch.float32, device=dev)
    topleft = torch.tensor([-cam.cx / cam.fx, -cam.cy / cam.fy], dtype=torch.float32).to(dev)
    bg = renderer.bg_rgb if renderer.bg else None
    ops.composite_sh_forward(k1["records"], renderer.sh_coeffs.data, start, end, ids, out, topleft, c2w, tile,
                             nth, ntw, 1.0 / cam.fx, 1.0 / cam.fy, H, W, C, renderer.T_thresh, bg_rgb=bg,
                             exact=renderer.exact_decisions)
    return out.view(H, W, 3), k1


@torch.no_grad()
def tile_sharded_render(renderer, c2w, camera_info, group=None, gather=True):
    """Render one frame with tile rows sharded over the ranks of `group` (forward only).
    Every rank returns the full [H,W,3] image when gather=True, else (its band [rows,W,3], row0)."""
    from . import ops

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    dev = renderer.mean.device
    cam, tile = camera_info, renderer.tile_size
    H, W = cam.h, cam.w
    nth = H // tile + (H % tile > 0)
    k1 = ops.project_cull_fused(
        renderer.mean.data, renderer.qvec.data, renderer.svec_before_activation.data,
        renderer.alpha_before_activation.data, renderer._svec_code, renderer._alpha_code,
        c2w.contiguous().float(), cam, renderer.frustum_culling_radius, renderer.skip_frustum_culling,
        renderer.tile_culling_radius, tile, cnt=None, want_records=True, want_activated=False)
    bands = balanced_bands(row_duplicate_counts(k1["tl"], k1["br"], nth), world)
    r0, r1 = bands[rank]
    img, _ = render_band(renderer, c2w, cam, r0, r1, k1=k1)
    y0, y1 = r0 * tile, min(r1 * tile, H)
    if world == 1:
        return img
    if not gather:
        return img[y0:y1], y0
    # equal-size padded bands so that one all_gather_into_tensor moves everything
    max_rows = max((min(b * tile, H) - a * tile) for a, b in bands)
    send = torch.zeros(max_rows, W, 3, dtype=torch.float32, device=dev)
    send[: y1 - y0] = img[y0:y1]
    recv = torch.empty(world * max_rows, W, 3, dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(recv, send, group=group)
    full = torch.empty(H, W, 3, dtype=torch.float32, device=dev)
    for k, (a, b) in enumerate(bands):
        ya, yb = a * tile, min(b * tile, H)
        full[ya:yb] = recv[k * max_rows: k * max_rows + (yb - ya)]
    return full
