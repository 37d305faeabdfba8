import sys, struct, subprocess, bisect, collections
out=sys.argv[1]
W=8192
raw=open(out+'.stk','rb').read()
ns=len(raw)//(8*(W+1))
maps=[]
for l in open(out+'.maps'):
    f=l.split()
    a,b=[int(x,16) for x in f[0].split('-')]
    maps.append((a,b,int(f[2],16),f[5] if len(f)>5 else '[anon]'))
exe=[m for m in maps if m[3].endswith(('turing_ref','turing_b200_batched','turing_b200_segments'))]
path=exe[0][3]; base=min(m[0] for m in exe); lo=base; hi=max(m[1] for m in exe)
o=subprocess.run(['nm','-C','--defined-only','-n',path],capture_output=True,text=True).stdout
t=[]
for l in o.splitlines():
    p=l.split(' ',2)
    if len(p)==3 and p[1] in 'tTwW': t.append((int(p[0],16),p[2]))
key=[x[0] for x in t]
pie = key[0] < 0x400000
def sym(pc):
    if not (lo<=pc<hi): return None
    rel=pc-base if pie else pc
    i=bisect.bisect_right(key,rel)-1
    return t[i][1] if i>=0 else None
REGIONS=[('searchMotionUni','me_uni'),('searchMotionBi','me_bi'),('subPelRefinement','me_uni/bi'),('fullPelMotionEstimation','me_uni'),('measurePuCost','pu_cost'),
         ('ReconstructInterBlock','tu_inter'),('ReconstructIntraBlock','tu_intra'),('predictIntraLuma','intra_sweep'),('reconstructIntraChroma','tu_intra_chroma'),('searchIntraChroma','intra_chroma'),
         ('TaskDeblock','deblock'),('TaskSao','sao'),('predictInter','predict_inter_final'),('Search<prediction_unit>','pu_other'),('searchIntraPartition','intra_partition_other'),
         ('reconstructInter','reconstruct_inter_other'),('Search<coding_unit>','cu_other'),('Search<coding_quadtree>','cqt_other'),('TaskEncodeSubstream','substream_other'),('TaskEncodeOutput','output'),('TaskEncodeInput','input')]
reg=collections.Counter(); leaf=collections.defaultdict(collections.Counter)
for i in range(ns):
    rec=struct.unpack_from('%dQ'%(W+1),raw,i*8*(W+1))
    pc=rec[0]; s=sym(pc)
    leafname = 'JIT' if s is None and not any(a<=pc<b and p.startswith('/') for a,b,_,p in maps) else (s or 'lib')
    if leafname.startswith('Rdoq::'): leafname='Rdoq::*'
    chain=[s] if s else []
    for wv in rec[1:]:
        x=sym(wv)
        if x: chain.append(x)
    r='unattributed'
    for name in chain:
        hit=[lab for pat,lab in REGIONS if pat in name.split('(')[0] or pat in name[:60]]
        if hit: r=hit[0]; break
    reg[r]+=1; leaf[r][leafname[:60]]+=1
print('samples',ns)
for r,v in reg.most_common():
    print(f'{100*v/ns:5.1f}%  {r:26s}', ', '.join(f'{k[:34]} {100*c/ns:.1f}' for k,c in leaf[r].most_common(4)))
