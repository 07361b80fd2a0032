"""K2 (binning) alone on the cfg-2 scene: CUDA-event time of the whole call and, with --kernels, the per-kernel
durations from the CUPTI activity trace (torch.profiler; warm caches, back-to-back launches -- unlike ncu's
serialised cold-cache replays).  GS3D_SORT=classic selects the round-1 three-kernel passes.

    python tools/bench_binning.py [cfg2|cfg5] [--kernels] [--save ids.pt]
"""
import sys
import collections
import torch
sys.path.insert(0, '/root/repo')
from gaussian_splatting_3d_b200 import ops, synthetic as S

args = [a for a in sys.argv[1:] if not a.startswith('--')]
cfg = args[0] if args else 'cfg2'
dev = 'cuda:0'
cam = S.make_camera(cfg); sc = S.make_scene(cfg, seed=0)
sc_d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in sc.items()}
def k1run():
    return ops.project_cull_fused(sc_d['mean'], sc_d['qvec'], sc_d['svec_before_activation'],
                                  sc_d['alpha_before_activation'], 1, 1, sc_d['c2w'], cam, 1.0, False, 6.0, 16)
k1 = k1run()
n = k1['n_dub']; nth, ntw = (cam.h + 15) // 16, (cam.w + 15) // 16
ids = torch.empty(n, dtype=torch.int32, device=dev)
st = torch.empty(nth * ntw, dtype=torch.int32, device=dev); en = torch.empty_like(st)
def run():
    ops.tile_culling_aabb_start_end(k1['tl'], k1['br'], ids, st, en, k1['depth'], nth, ntw, check_count=False)
for _ in range(5): run()
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print('binning ms', e0.elapsed_time(e1) / 20, 'n_dub', n)
if '--save' in sys.argv:
    path = sys.argv[sys.argv.index('--save') + 1]
    torch.save({'ids': ids.cpu(), 'start': st.cpu(), 'end': en.cpu()}, path)
def k1fast():
    return ops.project_cull_fused(sc_d['mean'], sc_d['qvec'], sc_d['svec_before_activation'],
                                  sc_d['alpha_before_activation'], 1, 1, sc_d['c2w'], cam, 1.0, False, 6.0, 16,
                                  want_activated=False, want_projection=False, sync_count=False)
if '--kernels' in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5): run(); k1fast()
        torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for ev in prof.events():
        if ev.device_type.name != 'CUDA':
            continue
        c = agg.setdefault(ev.name[:90], [0, 0.0])
        c[0] += 1; c[1] += ev.device_time
    tot = sum(v[1] for v in agg.values())
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'{t / 5:9.1f} us/call {c // 5:3d}x  {k}')
    print(f'{tot / 5:9.1f} us/call total kernel time')
for f, nm in ((k1run, 'K1 (all outputs, count read back)'), (k1fast, 'K1 (product path: records only, no sync)')):
    for _ in range(3): f()
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    print(nm, 'ms', e0.elapsed_time(e1) / 20)
