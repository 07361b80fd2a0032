import torch, time, sys
sys.path.insert(0,'/root/repo')
from gaussian_splatting_3d_b200 import ops, synthetic as S
dev='cuda:0'
cam=S.make_camera('cfg2'); sc=S.make_scene('cfg2',seed=0)
k1=ops.project_cull_fused(sc['mean'].to(dev), sc['qvec'].to(dev), sc['svec_before_activation'].to(dev), sc['alpha_before_activation'].to(dev),1,1,sc['c2w'].to(dev),cam,1.0,False,6.0,16)
n=k1['n_dub']; nth,ntw=(cam.h+15)//16,(cam.w+15)//16
ids=torch.empty(n,dtype=torch.int32,device=dev); st=torch.empty(nth*ntw,dtype=torch.int32,device=dev); en=torch.empty_like(st)
def run(): ops.tile_culling_aabb_start_end(k1['tl'],k1['br'],ids,st,en,k1['depth'],nth,ntw,check_count=False)
for _ in range(5): run()
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print('binning ms', e0.elapsed_time(e1)/20, 'n_dub', n)
def k1run(): return ops.project_cull_fused(sc_d['mean'], sc_d['qvec'], sc_d['svec_before_activation'], sc_d['alpha_before_activation'],1,1,sc_d['c2w'],cam,1.0,False,6.0,16)
sc_d={k:(v.to(dev) if torch.is_tensor(v) else v) for k,v in sc.items()}
for _ in range(3): k1run()
torch.cuda.synchronize(); e0.record()
for _ in range(20): k1run()
e1.record(); torch.cuda.synchronize()
print('K1 ms', e0.elapsed_time(e1)/20)
