import csv, collections, sys, subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]; iS=hdr.index('Source'); iE=hdr.index('Instructions Executed'); iN=hdr.index('# Samples')
iW=hdr.index('L1 Wavefronts Shared')
ops=collections.Counter(); samp=collections.Counter(); wf=collections.Counter(); tot=0; totw=0
body=[]
for r in rows[2:]:
    if len(r)<=iE: continue
    ins=r[iS].strip()
    try: e=int(r[iE]); n=int(r[iN]); w=int(r[iW])
    except: continue
    toks=ins.split()
    op=toks[1] if toks[0].startswith('@') else toks[0]
    full=op
    op=op.split('.')[0]
    if op in('LDS','STS','LDG','STG','RED','ATOMS','ATOMG','LDGSTS','UBLKCP','SYNCS'): op=full
    ops[op]+=e; samp[op]+=n; wf[op]+=w; tot+=e; totw+=w
    body.append((e,n,w,ins))
print('total warp-instr',tot,'smem wavefronts',totw)
for op,c in ops.most_common(40): print(f'{op:22s} {c/tot*100:6.2f}%  {c/1e6:8.1f}M  samples {samp[op]:6d}  wf {wf[op]/1e6:7.1f}M')
if len(sys.argv)>2:
    for e,n,w,ins in body:
        print(f'{e:10d} {n:6d} {w:9d}  {ins}')
