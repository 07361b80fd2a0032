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

    push=True (symmetric memory as well): the backward kernels accumulate into a PRIVATE buffer and mark
    the Gaussians they touch; exchange() launches one kernel that adds each marked row -- once, already
    summed over this rank's tiles and views -- into the RESULT buffer (the one `.grad` aliases) of every
    rank through the NVSwitch multicast address (multimem.red) or peer by peer, and sets the row's byte
    in every rank's union marks; only marked rows are ever cleared.  Two result buffers alternate, so ONE
    cross-rank barrier per step orders everything.  No NCCL call, no host round trip, U_own * 240 B leave
    each GPU per step (gs3d_rows_push_marked / gs3d_rows_zero_marked).  After exchange(), `.flat`,
    `.views` and the parameters' `.grad` refer to the buffer that holds this step's sum.

    pull=True: same buffers and marks, but the sum is formed INSIDE the NVSwitch: the ranks OR their marks
    into every rank's union marks, and each rank then reads the sum over all ranks of every n-th chunk
    of union rows with multimem.ld_reduce and broadcasts it with multimem.st (gs3d_marks_broadcast /
    gs3d_rows_pull_marked): each GPU ingests ~(1 + 1/n) * U rows instead of n * U -- the form that
    scales to 8 ranks, where the push form is bound by the NVLink packet rate of 16-byte reductions.

    sparse_reset=True (single GPU as well): zero() clears only the rows marked by the previous backward.

    sparse=True: the compositing backward marks the Gaussians whose gradient rows it writes; exchange()
    ORs the marks over the ranks (3 MB all-reduce), packs the union rows of all five blocks into one
    [U, 60] matrix (gs3d_rows_gather), all-reduces that -- U * 240 B instead of N * 236 B -- and
    unpacks the sum (gs3d_rows_scatter).  Untouched rows are zero on every rank and stay zero.
    """

    ORDER = ("sh_coeffs", "mean", "qvec", "svec_before_activation", "alpha_before_activation")

    def __init__(self, module, fused=False, group=None, use_multicast=True, sparse=False, push=False, pull=False,
                 sparse_reset=False):
        self.names = list(self.ORDER)
        self.params = [getattr(module, n) for n in self.names]
        # every block starts on a 16-byte boundary (the kernels use float4 accesses on qvec gradients and
        # vector reductions on SH rows; an unpadded layout is only aligned when N is a multiple of 4)
        total = sum((p.numel() + 3) // 4 * 4 for p in self.params)
        dev = self.params[0].device
        N = self.params[0].size(0)
        self.group = group
        multi = dist.is_initialized() and dist.get_world_size(group) > 1
        self.pull = bool(pull) and multi              # pull implies the push bookkeeping (marks, two result buffers)
        self.push = (bool(push) or self.pull) and multi
        self.fused = bool(fused) and multi and not self.push
        self.sparse = bool(sparse) and not self.fused and not self.push
        # sparse_reset (any world size, no exchange implied): zero() clears only the rows the previous
        # backward touched (marks from the compositing backward) instead of filling all 708 MB -- a view
        # touches a few percent of the Gaussians
        self.sparse_reset = bool(sparse_reset) and not self.fused and not self.push and dev.type == "cuda"
        self.touched = (torch.zeros(N, dtype=torch.uint8, device=dev)
                        if (self.sparse or self.push or self.sparse_reset) else None)
        # persistent 2-D gradient scratch (d loss / d mean2d, cov2d, alpha of the compositing backward): cleared
        # through the same marks as the leaf rows instead of three full-size fills per backward
        self.g2d = ([torch.zeros(N, 2, dtype=torch.float32, device=dev), torch.zeros(N, 4, dtype=torch.float32, device=dev),
                     torch.zeros(N, dtype=torch.float32, device=dev)]
                    if (self.touched is not None and not self.sparse and dev.type == "cuda") else None)
        self.last_union_rows = None
        self.handle = None
        self.local = self.local_views = self.union = self.union_handle = None
        def carve(buf):
            views, offsets, off = [], [], 0
            for p in self.params:
                views.append(buf[off:off + p.numel()].view_as(p))
                offsets.append(off)
                off += (p.numel() + 3) // 4 * 4
            return views, offsets

        self.peer_ptrs = None
        self.multicast_ptr = None
        self._res = None
        if self.fused or self.push:
            import torch.distributed._symmetric_memory as symm_mem

            grp = group if group is not None else dist.group.WORLD

            def symmetric(n, dtype):
                t = symm_mem.empty(n, dtype=dtype, device=dev)
                h = symm_mem.rendezvous(t, grp)
                t.zero_()
                return t, h

            self.flat, self.handle = symmetric(total, torch.float32)
            if self.push:
                # two result buffers used alternately: the one for step k+1 is reset during step k, so the
                # barrier that ends step k's exchange also orders "reset" before "peers push into it"
                self._res = []
                for b in range(2):
                    buf, h = (self.flat, self.handle) if b == 0 else symmetric(total, torch.float32)
                    uni, uh = symmetric(N, torch.uint8)
                    mc = int(h.multicast_ptr) if use_multicast else 0
                    umc = int(uh.multicast_ptr) if use_multicast else 0
                    self._res.append(dict(flat=buf, handle=h, views=carve(buf)[0], union=uni,
                                          peer_ptrs=[int(p) for p in h.buffer_ptrs],
                                          union_ptrs=[int(p) for p in uh.buffer_ptrs], multicast=mc if mc else None,
                                          union_multicast=umc if umc else None))
                self._cur = 0
                self.union = self._res[0]["union"]
                if self.pull:  # peers read this rank's private rows: symmetric (multicast-mapped) as well
                    self.local, lh = symmetric(total, torch.float32)
                    lmc = int(lh.multicast_ptr) if use_multicast else 0
                    self._local_peer_ptrs = [int(p) for p in lh.buffer_ptrs]
                    self._local_multicast = lmc if lmc else None
                    self._rank = dist.get_rank(grp)
                else:
                    self.local = torch.zeros(total, dtype=torch.float32, device=dev)
        else:
            self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views, self.offsets = carve(self.flat)
        for p, v in zip(self.params, self.views):
            p.grad = v
        if self.local is not None:
            self.local_views = carve(self.local)[0]
        self.n_sh = self.params[0].numel()
        if self.fused or self.push:
            # the SH block starts at offset 0 of the symmetric buffer on every rank
            self.peer_ptrs = [int(p) for p in self.handle.buffer_ptrs]
            mc = int(self.handle.multicast_ptr) if use_multicast else 0
            self.multicast_ptr = mc if mc else None
        self.module = None
        self._dirty = False
        self._backwards_since_zero = 0
        # side stream for the per-step resets and the mark broadcast: they overlap the forward kernels and
        # the projection backward instead of sitting on the critical path between two barriers
        self._aux = torch.cuda.Stream(device=dev) if (self.push and dev.type == "cuda") else None
        self._aux_ev = None

    def _on_aux(self, fn):
        """Run fn on the side stream after everything enqueued so far; returns the completion event."""
        cur = torch.cuda.current_stream(self.flat.device)
        self._aux.wait_stream(cur)
        with torch.cuda.stream(self._aux):
            fn()
            ev = torch.cuda.Event()
            ev.record(self._aux)
        return ev

    def _after_composite_backward(self):
        """Called by the backward right after the compositing-backward launch (marks are final once it
        completes): the mark broadcast runs beside the projection backward."""
        from . import ops

        cur = self._res[self._cur]
        self._marks_ev = self._on_aux(lambda: ops.marks_broadcast(self.touched, cur["union_ptrs"],
                                                                      cur["union_multicast"]))

    def attach(self, renderer):
        bufs = dict(zip(self.names, self.local_views if self.push else self.views))
        if self.pull and self._aux is not None:
            bufs["after_composite_backward"] = self._after_composite_backward
        if self.fused:
            bufs["sh_peer_ptrs"] = self.peer_ptrs
            bufs["sh_multicast_ptr"] = self.multicast_ptr
        if self.touched is not None:
            bufs["touched"] = self.touched
        if self.g2d is not None:
            bufs["g_mean2d"], bufs["g_cov2d"], bufs["g_alpha2d"] = self.g2d
        renderer.grad_buffers = bufs
        self.module = renderer
        return self

    def detach(self):
        if self.module is not None:
            self.module.grad_buffers = None
            self.module = None

    def _check_attached(self):
        """The renderer drops `grad_buffers` when it replaces its parameters (split / prune / load):
        this object then still aliases the OLD tensors and must be rebuilt."""
        m = self.module
        if m is not None and (m.grad_buffers is None or any(getattr(m, n) is not p
                                                             for n, p in zip(self.names, self.params))):
            raise RuntimeError("FlatGradients: the renderer's parameters were replaced (adaptive density control "
                               "or load) after attach(); build a new FlatGradients(renderer).attach(renderer)")

    def _barrier(self):
        if self.handle is not None:
            self.handle.barrier(channel=0)
        elif dist.is_initialized():
            dist.barrier(group=self.group)

    def zero(self):
        self._check_attached()
        self._backwards_since_zero = 0
        if self.push and self.module is not None:
            from . import ops

            # (i) this rank's private rows written last step (own marks) and (ii) the OTHER result buffer --
            # last step's sum, dead once zero() is called, pushed into again only after this step's final
            # barrier -- are reset on the side stream while the forward kernels run
            cur, nxt = self._res[self._cur], self._res[self._cur ^ 1]

            def resets():
                ops.rows_zero_marked(self.touched, list(self.local_views) + (self.g2d or []), clear_marks=True)
                ops.rows_zero_marked(nxt["union"], nxt["views"], clear_marks=True)

            if self._aux is not None:
                self._aux_ev = self._on_aux(resets)
            else:
                resets()
            self._marks_ev = None
            self.flat, self.views, self.union = cur["flat"], cur["views"], cur["union"]
        elif self.sparse_reset and self.module is not None:
            from . import ops

            ops.rows_zero_marked(self.touched, list(self.views) + (self.g2d or []), clear_marks=True)
        else:
            self.flat.zero_()
            if self.touched is not None:
                self.touched.zero_()
        for p, v in zip(self.params, self.views):
            p.grad = v  # optimisers / zero_grad(set_to_none) may have dropped the alias
        if self.fused:
            self._barrier()  # nobody may add into a buffer that is not zeroed yet

    def backward_into(self, loss):
        """Run backward for `loss`; the kernels add the leaf gradients into the flat buffer."""
        self._check_attached()
        if self._aux_ev is not None:  # the private buffer must be clean before the backward adds into it
            torch.cuda.current_stream(self.flat.device).wait_event(self._aux_ev)
            self._aux_ev = None
        if self.module is None:
            loss.backward()  # plain autograd accumulation into the aliased .grad views
        else:
            if self.g2d is not None and self._backwards_since_zero > 0:
                # a second view in the same step: the persistent 2-D gradient scratch still holds the previous
                # view's d loss / d (mean2d, cov2d, alpha) -- they belong to THAT view's projection and were consumed
                # by its projection backward.  Clear the marked rows (the marks themselves stay: they describe
                # the step's union of touched leaf rows).
                from . import ops

                ops.rows_zero_marked(self.touched, self.g2d, clear_marks=False)
            self._backwards_since_zero += 1
            # the attached renderer's forward made a fresh one-element leaf its only differentiable input
            # (gs/renderer.py splat_sh): backward() writes the leaf gradients into this object's buffers itself
            anchor = (self.module._state or {}).get("anchor") if hasattr(self.module, "_state") else None
            torch.autograd.grad(loss, [anchor] if anchor is not None else self.params)

    def exchange(self, average=False):
        """Make every rank's buffer hold the sum over ranks."""
        if not (dist.is_initialized() and dist.get_world_size(self.group) > 1):
            return self.flat
        if self.fused and self.module is not None:
            self._barrier()  # every rank's in-kernel reductions into this buffer have landed
            dist.all_reduce(self.flat[self.n_sh:], op=dist.ReduceOp.SUM, group=self.group)
        elif self.push and self.module is not None:
            from . import ops

            cur = self._res[self._cur]
            stream = torch.cuda.current_stream(self.flat.device)
            if self._aux_ev is not None:  # (no backward ran this step)
                stream.wait_event(self._aux_ev)
                self._aux_ev = None
            if self.pull:
                if self._marks_ev is not None:  # broadcast already issued beside the projection backward
                    stream.wait_event(self._marks_ev)
                    self._marks_ev = None
                else:
                    ops.marks_broadcast(self.touched, cur["union_ptrs"], cur["union_multicast"])
                cur["handle"].barrier(channel=0)  # union marks complete, every rank's private rows final
                widths = [v.numel() // v.size(0) for v in self.views]
                ops.rows_pull_marked(cur["union"], widths, self.offsets, self._local_peer_ptrs, cur["peer_ptrs"],
                                     self._rank, self._local_multicast, cur["multicast"])
            else:
                ops.rows_push_marked(self.touched, self.local_views, self.offsets, cur["peer_ptrs"],
                                     cur["union_ptrs"], cur["multicast"])
            cur["handle"].barrier(channel=0)  # every rank's rows have landed; every 'next' buffer is clean
            self._cur ^= 1
        elif self.sparse and self.module is not None:
            self._exchange_sparse()
        else:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.group)
        if average:
            self.flat.div_(dist.get_world_size(self.group))
        return self.flat

    def _exchange_sparse(self):
        from . import ops

        marks = self.touched
        dist.all_reduce(marks, op=dist.ReduceOp.MAX, group=self.group)   # OR of the ranks' marks
        idx = union_rows(marks)                                           # same list on every rank
        self.last_union_rows = int(idx.numel())
        if idx.numel() == 0:
            return
        width = sum(v.numel() // v.size(0) for v in self.views)
        packed = torch.empty(idx.numel(), (width + 3) // 4 * 4, dtype=torch.float32, device=self.flat.device)
        ops.rows_gather(self.views, idx, packed)
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=self.group)
        ops.rows_scatter(self.views, idx, packed)

    def all_reduce(self, group=None, average=False):
        return self.exchange(average=average)


def union_rows(marks):
    """int32 indices of the non-zero entries of the (already OR-reduced) per-Gaussian marks, ascending;
    identical on every rank because the marks are."""
    return torch.nonzero(marks, as_tuple=False).view(-1).to(torch.int32)


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
    Each rank contributes only what it accumulated since the last sync.

    The baseline of the last sync is dropped whenever the renderer replaces its ADC buffers
    (adaptive_control -> _reset_adc_buffers after a split / prune / alpha reset): a baseline that
    does not belong to the current buffers -- other shape, or the buffers were re-zeroed -- restarts
    from zero.  With split_type 'mean_grad' and reduction 'mean' the statistic is ||mean.grad|| of the
    ALREADY exchanged gradient, identical on every rank, so it is not summed again (that would inflate
    it by the world size); only the fused / 2-D statistic is rank-local."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return
    base = getattr(renderer, "_adc_synced", None)
    epoch = getattr(renderer, "_adc_epoch", 0)
    if (base is None or base[2] != epoch or base[0].shape != renderer.grad_mean.shape
            or base[1].shape != renderer.cnt.shape):
        base = (torch.zeros_like(renderer.grad_mean), torch.zeros_like(renderer.cnt), epoch)
    gm0, cnt0, _ = base
    d_cnt = renderer.cnt - cnt0
    dist.all_reduce(d_cnt, op=dist.ReduceOp.SUM, group=group)
    renderer.cnt = cnt0 + d_cnt
    rank_local = getattr(renderer, "split_type", "2d_mean_grad") != "mean_grad"
    if renderer.split_reduction == "max":
        gm = renderer.grad_mean.clone()
        if rank_local:
            dist.all_reduce(gm, op=dist.ReduceOp.MAX, group=group)
    elif rank_local:
        d_gm = renderer.grad_mean - gm0
        dist.all_reduce(d_gm, op=dist.ReduceOp.SUM, group=group)
        gm = gm0 + d_gm
    else:
        gm = renderer.grad_mean.clone()
    renderer.grad_mean = gm
    renderer._adc_synced = (gm.clone(), renderer.cnt.clone(), epoch)


# ---------------------------------------------------------------- tile-sharded rendering


def row_duplicate_counts(aabb_topleft, aabb_bottomright, n_tiles_h):
    """Duplicates per tile row from the per-Gaussian rects (int64 [n_tiles_h]); integer arithmetic only,
    so every rank computes the same numbers.  CUDA tensors: gs3d_row_duplicate_counts (per-block
    difference arrays); CPU tensors (host-logic tests): the same difference array in torch."""
    if aabb_topleft.is_cuda:
        from . import ops

        return ops.row_duplicate_counts(aabb_topleft, aabb_bottomright, n_tiles_h)
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
def render_band(renderer, c2w, camera_info, row_begin=None, row_end=None, k1=None, out=None):
    """Forward-render tile rows [row_begin, row_end) of one frame (None = all rows).  Returns the
    full-size [H,W,3] buffer in which only the band's rows are DEFINED (the rest is uninitialised memory), plus
    the K1 outputs."""
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
    # clip to the band and keep only the Gaussians that still cover a tile: the band's depth sort and
    # emission then run over ~N/world Gaussians; ids come back as original indices through `index`
    tl, br, depth_b, index, n_band = ops.clip_rects_to_rows(k1["tl"], k1["br"], k1["depth"], row_begin, row_end)
    ids = torch.empty(n_band, dtype=torch.int32, device=dev)
    start = torch.empty(nth * ntw, dtype=torch.int32, device=dev)
    end = torch.empty(nth * ntw, dtype=torch.int32, device=dev)
    ops.tile_culling_aabb_start_end(tl, br, ids, start, end, depth_b, nth, ntw, check_count=False)
    if n_band:
        ids = torch.index_select(index, 0, ids)
    # only the band's rows are ever read back: empty tiles of the band keep the zero fill, the rest of the
    # full-size buffer stays uninitialised (a 4K frame is 99.5 MB: no full fill per band)
    # `out` may be another rank's frame buffer mapped into this process (symmetric memory): the compositing kernel
    # then stores the band's pixels straight over NVLink
    if out is None:
        out = torch.empty(H * W * 3, dtype=torch.float32, device=dev)
    out[row_begin * tile * W * 3: min(row_end * tile, H) * W * 3].zero_()
    topleft = renderer._topleft(cam) if hasattr(renderer, "_topleft") else \
        torch.tensor([-cam.cx / cam.fx, -cam.cy / cam.fy], dtype=torch.float32).to(dev)
    bg = renderer.bg_rgb if renderer.bg else None
    ops.composite_sh_forward(k1["records"], renderer.sh_coeffs.data, start, end, ids, out, topleft, c2w, tile,
                             nth, ntw, 1.0 / cam.fx, 1.0 / cam.fy, H, W, C, renderer.T_thresh, bg_rgb=bg,
                             exact=renderer.exact_decisions)
    return out.view(H, W, 3), k1


class SharedFrame:
    """A frame buffer on `root` that every rank of the group can store into: torch symmetric memory, each
    rank holds a tensor view of the ROOT's buffer (peer mapping over NVLink).  tile_sharded_render(frame=...)
    composites every band straight into it -- no gather step, no staging copies; two symmetric-memory barriers
    per frame order 'root is done with the previous frame' -> stores -> 'all bands have landed'."""

    def __init__(self, camera_info, device, group=None, root=0):
        import torch.distributed._symmetric_memory as symm_mem

        grp = group if group is not None else dist.group.WORLD
        n = camera_info.h * camera_info.w * 3
        self.local = symm_mem.empty(n, dtype=torch.float32, device=device)
        self.handle = symm_mem.rendezvous(self.local, grp)
        self.root = root
        self.rank = dist.get_rank(group)
        self.shape = (camera_info.h, camera_info.w, 3)
        self.root_view = self.local if self.rank == root else self.handle.get_buffer(root, (n,), torch.float32)

    def image(self):
        """The finished frame ([H,W,3]) on the root, None elsewhere."""
        return self.local.view(self.shape) if self.rank == self.root else None


@torch.no_grad()
def tile_sharded_render(renderer, c2w, camera_info, group=None, gather=True, frame=None):
    """Render one frame with tile rows sharded over the ranks of `group` (forward only).
    frame=None: every rank returns the full [H,W,3] image when gather=True (one padded all_gather), else (its band
    [rows,W,3], row0).  frame=SharedFrame: the bands are composited straight into the root's frame buffer over
    NVLink; returns frame.image() (the full frame on the root, None elsewhere)."""
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
        renderer.tile_culling_radius, tile, cnt=None, want_records=True, want_activated=False,
        sync_count=False, want_projection=False)  # (the whole-frame duplicate count is not needed here)
    bands = balanced_bands(row_duplicate_counts(k1["tl"], k1["br"], nth), world)
    r0, r1 = bands[rank]
    if frame is not None and world > 1:
        frame.handle.barrier(channel=0)  # the root has consumed the previous frame (stream order on the root)
        render_band(renderer, c2w, cam, r0, r1, k1=k1, out=frame.root_view)
        frame.handle.barrier(channel=0)  # every band has landed in the root's buffer
        return frame.image()
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
