"""Per-kernel CUPTI times of the eager training step of a config (default cfg3: forward, L2 loss, backward, ADC
accumulation, fused Adam).  python tools/profile_step.py [cfg3|cfg2]"""
import collections
import sys
import torch
sys.path.insert(0, '/root/repo')
from gaussian_splatting_3d_b200 import synthetic as S
from torch.profiler import profile, ProfilerActivity
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
dev = 'cuda:0'
cam = S.make_camera(name); sc = S.make_scene(name, seed=0)
r = S.renderer_from_scene(sc, S.make_cfg(device=dev, sh_order=sc['C'], warm_up=0, fused_adam=True, adam_single_step=True))
r.train(); r.fuse_adc = True
c2w = sc['c2w'].to(dev); tgt = S.make_target(cam, 0).to(dev)
state = {'opt': r.get_optimizer(0), 'e': 0}
def step():
    o = r(c2w, cam); loss = ((o - tgt) ** 2).mean()
    state['opt'].zero_grad(); loss.backward(); state['opt'].step()
    r.adaptive_control(state['e']); state['opt'] = r.get_optimizer(state['e']); state['e'] += 1
for _ in range(3): step()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): step()
e1.record(); torch.cuda.synchronize()
print(name, 'step ms', e0.elapsed_time(e1) / 20, 'n_dub', r.total_dub_gaussians)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
agg = collections.OrderedDict()
for ev in prof.events():
    if ev.device_type.name != 'CUDA': continue
    c = agg.setdefault(ev.name[:70], [0, 0.0]); c[0] += 1; c[1] += ev.device_time
tot = sum(v[1] for v in agg.values())
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print(f'{t / 5:9.1f} us/step {c / 5:5.1f}x  {k}')
print(f'{tot / 5:9.1f} us/step total kernel time')
