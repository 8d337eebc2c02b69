"""Does the dead-path rule (with its guard) change a bit anywhere?  On the CPU, with the oracle's test switch, over full
1080p frames of the bench workload: every sample index below 20 000 at which some pixel's key draws a random number of
exactly 1.0 in some dimension, plus ten random sample indices below 2^20 — reference loop vs retiring loop, accumulators
compared bit for bit (NaN payloads included).  Output: profiles/r2_dead_path_scan.txt.
usage: python tests/checkers/retire_scan.py"""
import sys, os, json, time
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, 'oracle'))
import numpy as np, bench, oracle as om
P=np.array([0x6a09e667,0xbb67ae84,0x3c6ef372,0xa54ff539,0x510e527f,0x9b05688a,0x1f83d9ab,0x5be0cd18,0xcbbb9d5c,0x629a2929,0x91590159,0x452fecd8,0x67332667,0x8eb44a86,0xdb0c2e0b,0x47b5481d,0xae5f9155,0xcf6c85d1,0x2f73477d,0x6d1826ca,0x8b43d455,0xe360b595,0x1c456002,0x6f196330,0xd94ebeaf,0x9cc4a611,0x261dc1f2,0x5815a7bd,0x70b7ed67,0xa1513c68,0x44f93634,0x720dcdfc],np.uint64)
world, cfg, seeds, _, label, scene, sky = bench.load_workload('breaktime')
offsets=np.unique(seeds[:,1]).astype(np.uint64)
# sample indices below N whose key draws a 1.0 in some dimension, for some offset present in the frame
N=20000
special={}
for n0 in range(0,N,2000):
    n=np.arange(n0,n0+2000,dtype=np.uint64)
    key=(n[:,None]+offsets[None,:])&0xFFFFFFFF            # [n, offsets]
    q=(key[:,:,None]*P[None,None,1:])&0xFFFFFFFF
    hit=(q>=0xFFFFFF80)
    for i,j,d in zip(*np.nonzero(hit)):
        special.setdefault(int(n[i]),[]).append((int(offsets[j]),int(d)+1))
print('sample indices < %d with a 1.0 somewhere: %s'%(N, {k:v for k,v in sorted(special.items())}), flush=True)
osc=om.OracleScene(world, sky)
rs=np.random.default_rng(11)
todo=sorted(special)[:12]+sorted(int(x) for x in rs.integers(0,1<<20,10))
res=[]
for k in todo:
    s=seeds.copy(); s[:,0]+=np.uint32(k)
    om.set_retire_dead_paths(False); a,_,ca,_=om.trace(cfg,osc,s.copy(),1)
    om.set_retire_dead_paths(True); b,_,cb,_=om.trace(cfg,osc,s.copy(),1)
    om.set_retire_dead_paths(False)
    same=bool((a.view(np.uint32)==b.view(np.uint32)).all())
    nan=int((~np.isfinite(a[:,:3]).all(axis=1)).sum())
    res.append((k, k in special, nan, same, cb['nearest_rays']/ca['nearest_rays']))
    print(res[-1], flush=True)
print('all identical:', all(r[3] for r in res))
