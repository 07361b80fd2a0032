"""Per-view cost of the 8 weak-scaling poses of bench.py on ONE GPU (is the per-rank spread at 8 GPUs the views
or the GPUs?).  python tools/view_costs.py"""
import sys
import torch
sys.path.insert(0, '/root/repo')
import bench
from gaussian_splatting_3d_b200 import synthetic as S, parallel as P
dev = 'cuda:0'
cam = S.make_camera('cfg2'); sc = S.make_scene('cfg2', seed=0)
r = S.renderer_from_scene(sc, S.make_cfg(device=dev, sh_order=sc['C'])); r.train()
flat = P.FlatGradients(r, sparse_reset=True).attach(r)
views = bench.views_for(8, torch)
tgt = S.make_target(cam, 0).to(dev)
r.static_capacity = 14 << 20
for i, v in enumerate(views):
    c2w = v.to(dev)
    def step():
        flat.zero(); out = r(c2w, cam); flat.backward_into(((out - tgt) ** 2).mean())
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    print(f'view {i}: {e0.elapsed_time(e1) / 10:.3f} ms/step  n_dub {int(r.total_dub_gaussians)}  overflow {r.overflowed()}')
import collections
from torch.profiler import profile, ProfilerActivity
for i in (2, 3):
    c2w = views[i].to(dev)
    def step():
        flat.zero(); out = r(c2w, cam); flat.backward_into(((out - tgt) ** 2).mean())
    for _ in range(2): step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3): step()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type.name != 'CUDA': continue
        c = agg.setdefault(ev.name[:60], [0, 0.0]); c[0] += 1; c[1] += ev.device_time
    print(f'--- view {i}')
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:6]:
        print(f'{t / 3:9.1f} us/step {c // 3:3d}x  {k}')
    # tile list statistics
    st, en = r._state['start'], r._state['end']
    ln = (en - st).clamp_min(0).float()
    print('tiles', ln.numel(), 'mean len', float(ln.mean()), 'max', float(ln.max()), 'p99', float(ln.kthvalue(int(0.99 * ln.numel())).values))
