"""A whole training step -- gradient reset, forward (cull + project + bin + composite), loss, backward to the
leaf parameters -- captured once in three CUDA graphs ([reset + forward], [loss], [backward]) and replayed with
three launches per step.  The loss graph is separate so that the step's result (loss + overflow flag, 8 bytes) can
be copied to the host as soon as it exists: a caller that reads the loss every step gets it while the backward
kernels (more than half of the step) still run, and has the next step enqueued before the GPU goes idle.

The reference's step has three host synchronisations inside the forward alone (boolean-mask `nonzero`,
`.item()` in gs/culling.py:33-35, the D2H memcpy + cudaFree in aabb_culling.h:204-259) and ~250 kernel
launches; this repo's eager step has one 8-byte read-back (the duplicate count) and ~40 launches.  With
`SHRenderer.static_capacity` set, the count never leaves the device (gs3d_tile_culling_aabb_start_end_capacity),
every launch is sized by the capacity, and the step becomes a static launch sequence: exactly what a CUDA
graph needs.  The only per-step host work left is copying the pose and the target image into the static input
buffers and (optionally) reading the loss back.

    flat = parallel.FlatGradients(renderer, sparse_reset=True).attach(renderer)
    step = graph.GraphedStep(renderer, flat, camera_info)        # captures on first use
    loss = step(c2w, target)                                     # tensors on the host (pinned) or the device
    step.check()                                                 # raises if the capacity overflowed (syncs)

The captured sequence is the SAME Python code path as the eager step (SHRenderer.forward, the L2 loss,
FlatGradients.backward_into), so results are identical to it; tests/test_gpu_graph.py compares them.
"""
import torch


class GraphedStep:
    def __init__(self, renderer, flat, camera_info, capacity=None, loss_fn=None, margin=1.25, warmup=3):
        """capacity: duplicates the tile lists may hold; default = margin x the count of one eager forward of
        the first call's view (rounded up to 64 K).  loss_fn(out, target) -> scalar; default mean squared error."""
        self.r, self.flat, self.cam = renderer, flat, camera_info
        self.capacity = capacity
        self.margin = margin
        self.warmup = warmup
        self.loss_fn = loss_fn or (lambda out, tgt: ((out - tgt) ** 2).mean())
        self.g_fwd = self.g_loss = self.g_bwd = None  # (the target image is first needed by the loss graph)
        dev = renderer.mean.device
        self.c2w = torch.zeros(3, 4, dtype=torch.float32, device=dev)
        self.target = torch.zeros(camera_info.h, camera_info.w, 3, dtype=torch.float32, device=dev)
        self.loss = None
        self.status = None  # [loss, overflow flag] packed for one 8-byte read-back
        self._status_host = torch.zeros(2, dtype=torch.float32).pin_memory()
        self._status_ready = torch.cuda.Event()
        self._copy_stream = torch.cuda.Stream(device=dev)

    # ------------------------------------------------------------------ capture
    def _forward(self):
        self.flat.zero()
        return self.r(self.c2w, self.cam)

    def _backward(self, out):
        loss = self.loss_fn(out, self.target)
        self.flat.backward_into(loss)
        return loss

    def _body(self):
        return self._backward(self._forward())

    def _capture(self):
        r = self.r
        dev = r.mean.device
        if self.capacity is None:
            r.static_capacity = None
            with torch.no_grad():
                r(self.c2w, self.cam)  # one eager forward to learn the scale of the duplicate count
            n = int(r.total_dub_gaussians)
            self.capacity = max(1 << 16, (int(n * self.margin) + 65535) // 65536 * 65536)
        r.static_capacity = int(self.capacity)
        # warm-up on a side stream (allocator, autograd, lazy module state), as torch.cuda.graph requires
        s = torch.cuda.Stream(device=dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(self.warmup):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        # Three graphs sharing one memory pool: [reset + forward], [loss], [backward].  The target image is first
        # read by the loss, so its host->device copy (13 MB at cfg 2) overlaps the forward graph; the loss and the
        # overflow flag are final before the backward starts, so their read-back overlaps the backward graph.
        self.g_fwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fwd):
            out = self._forward()
        self.g_loss = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_loss, pool=self.g_fwd.pool()):
            loss = self.loss_fn(out, self.target)
            self.loss = loss.detach()
            self.status = torch.stack([self.loss.reshape(()), r._overflow.reshape(()).to(torch.float32)])
        self.g_bwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_bwd, pool=self.g_fwd.pool()):
            self.flat.backward_into(loss)
        self.out = out
        # the capture leaves the parameters' .grad aliased to the flat buffer's views, as the eager step does

    # ------------------------------------------------------------------ replay
    def __call__(self, c2w, target, read_loss=False):
        """One step.  c2w [3,4], target [H,W,3]: host (ideally pinned) or device tensors.  Returns the loss as a
        device scalar (read_loss=False) or as a Python float together with the overflow check (read_loss=True)."""
        if self.g_fwd is None:
            self.c2w.copy_(c2w)
            self._capture()
        dev = self.r.mean.device
        main = torch.cuda.current_stream(dev)
        # the pose is read by the first kernel: same stream.  The 13 MB target is only read by the loss:
        # its copy runs beside the previous replay's tail / this replay's forward on the copy stream.
        self.c2w.copy_(c2w, non_blocking=True)
        side = target.device != self.target.device
        if side:
            self._copy_stream.wait_stream(main)  # the previous step's loss has read the static target
            with torch.cuda.stream(self._copy_stream):
                self.target.copy_(target, non_blocking=True)
        elif target.data_ptr() != self.target.data_ptr():
            self.target.copy_(target, non_blocking=True)
        self.g_fwd.replay()
        if side:
            main.wait_stream(self._copy_stream)
        self.g_loss.replay()
        if read_loss:  # 8 bytes to pinned memory, enqueued BEFORE the backward graph
            self._status_host.copy_(self.status, non_blocking=True)
            self._status_ready.record(main)
        self.g_bwd.replay()
        if read_loss:
            self._status_ready.synchronize()  # the loss is on the host; the backward is still running
            st = self._status_host.tolist()
            if st[1] != 0.0:
                raise RuntimeError(f"GraphedStep: the duplicate count exceeded the static capacity {self.capacity}; "
                                   "re-create the step with a larger capacity")
            return st[0]
        return self.loss

    def check(self):
        if self.status is not None and float(self.status[1].item()) != 0.0:
            raise RuntimeError(f"GraphedStep: the duplicate count exceeded the static capacity {self.capacity}")
