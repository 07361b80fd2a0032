"""Three eager cfg-2 training steps (reset, K1, K2, K3, L2 loss, K4a, K4b) for profiler captures; the last one is
bracketed by cudaProfilerStart / cudaProfilerStop (process-wide: the backward kernels are launched from autograd's
own thread, which a thread-local NVTX range would miss):

    ncu --set full --clock-control none --profile-from-start off -o rep python tools/one_step.py
"""
import sys
import torch
sys.path.insert(0, '/root/repo')
from gaussian_splatting_3d_b200 import synthetic as S, parallel as P

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
dev = 'cuda:0'
cam = S.make_camera(name); sc = S.make_scene(name, seed=0)
r = S.renderer_from_scene(sc, S.make_cfg(device=dev, sh_order=sc['C'])); r.train()
flat = P.FlatGradients(r, sparse_reset=True).attach(r)
c2w = sc['c2w'].to(dev); tgt = S.make_target(cam, 0).to(dev)
def step():
    flat.zero(); out = r(c2w, cam); flat.backward_into(((out - tgt) ** 2).mean())
for i in range(3):
    if i == 2:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    step()
    if i == 2:
        torch.cuda.synchronize(); torch.cuda.profiler.stop()
print('n_dub', r.total_dub_gaussians)
