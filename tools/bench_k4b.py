"""K4b (projection backward) alone on the cfg-2 scene with three row filters: the frustum mask (dense), the
compositing backward's touched marks through the one-thread-per-Gaussian kernel, and through the compacting kernel.

    python tools/bench_k4b.py
"""
import sys, torch
sys.path.insert(0, '/root/repo')
from gaussian_splatting_3d_b200 import synthetic as S, parallel as P, ops
dev = 'cuda:0'
cam = S.make_camera('cfg2'); sc = S.make_scene('cfg2', seed=0)
r = S.renderer_from_scene(sc, S.make_cfg(device=dev, sh_order=sc['C'])); r.train()
flat = P.FlatGradients(r, sparse_reset=True).attach(r)
c2w = sc['c2w'].to(dev); tgt = S.make_target(cam, 0).to(dev)
calls = []
orig = ops.project_backward_fused
def spy(mask, *a, **k):
    calls.append((mask.dtype, int(mask.count_nonzero()), mask.numel()))
    return orig(mask, *a, **k)
ops.project_backward_fused = spy
for i in range(2):
    flat.zero(); out = r(c2w, cam); flat.backward_into(((out - tgt) ** 2).mean())
torch.cuda.synchronize()
print(calls, 'touched now', int(flat.touched.count_nonzero()), 'mask', int(r.frustum_culling_mask.count_nonzero()))
ops.project_backward_fused = orig
import functools
st = r._state
bufs = r.grad_buffers
N = r.mean.shape[0]
g2 = [bufs['g_mean2d'], bufs['g_cov2d'], bufs['g_alpha2d']]
leaf = (bufs['mean'], bufs['qvec'], bufs['svec_before_activation'], bufs['alpha_before_activation'])
def run(maskt, sparse=False):
    ops.project_backward_fused(maskt, r.mean.data, r.qvec.data, r.svec_before_activation.data, r.alpha_before_activation.data,
                               1, 1, c2w, True, g2[0], g2[1], g2[2], out=leaf, accumulate=True, sparse_filter=sparse)
for nm, m, sp in (('frustum mask', r.frustum_culling_mask, False), ('touched marks', flat.touched, False), ('touched marks, compacting kernel', flat.touched, True), ('all zero', torch.zeros_like(flat.touched), False)):
    for _ in range(3): run(m, sp)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): run(m, sp)
    e1.record(); torch.cuda.synchronize()
    print(nm, 'K4b ms', e0.elapsed_time(e1) / 20)
