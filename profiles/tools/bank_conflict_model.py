"""Bank-conflict model of the K1 v10 tile gather (round 2): cell-sorted particles of one supercell, the lane maps "pair" (lane l owns
particles 2l, 2l+1: the kernel) and "split" (l, l+32), tile pitches NY x NZ; prints modelled shared-memory wavefronts per LDS for
Ex..Bz and the total per 32 particles.  Reproduces the ncu source-page numbers of profiles/r02_k1_ablation.md section 2."""
import numpy as np, sys
rng=np.random.default_rng(1)
def simulate(NY, order='zyx', ppc=8, nsc=200, W=2, drift=0.0, NZ=8, perm=None, lane_map='pair'):
    # supercell 4x4x4 cells at tile-local cells 2..5
    GT=[(1,0,0),(0,1,0),(0,0,1),(0,1,1),(1,0,1),(1,1,0)]
    SX=NY*NZ
    tot=np.zeros(6); cnt=0
    for sc in range(nsc):
        n=rng.poisson(ppc*64)
        pos=rng.random((n,3))*4+2.0   # tile coords of centre line q in [2,6)  (cell c => q in [c, c+1))
        # hmm: centre-line coordinate q: anchor ic=floor(q); cell = ic (array idx)  ok
        cell=np.floor(pos).astype(int)
        if order=='zyx': key=(cell[:,0]*8+cell[:,1])*8+cell[:,2]
        elif order=='zxy': key=(cell[:,1]*8+cell[:,0])*8+cell[:,2]
        idx=np.argsort(key,kind='stable'); pos=pos[idx]
        pos=pos+rng.normal(0,drift,(n,3)) if drift>0 else pos
        ic=np.floor(pos).astype(int); r=pos-ic
        iv=np.floor(pos-0.5).astype(int)
        CH=32*W
        cnt+=W*((n+CH-1)//CH)
        for c0 in range(0,n,CH):
            sl=slice(c0,min(n,c0+CH))
            for j in range(W):
                if lane_map=='pair': sel=np.arange(c0+j,min(n,c0+CH),W)
                else: sel=np.arange(c0+j*32,min(n,c0+j*32+32))
                if len(sel)==0: continue
                for c,(gx,gy,gz) in enumerate(GT):
                    ax=(iv if gx else ic)[sel,0]; ay=(iv if gy else ic)[sel,1]; az=(iv if gz else ic)[sel,2]
                    addr=ax*SX+ay*NZ+az
                    if perm is not None: addr=perm(ax,ay,az)
                    ua=np.unique(addr)
                    banks=ua%32
                    wf=np.bincount(banks,minlength=32).max()
                    tot[c]+=wf
    return tot/ cnt
for NY in (9,10,11,12):
    for order in ('zyx','zxy'):
        for lm in ('pair','split'):
            r=simulate(NY,order,lane_map=lm)
            print(NY,order,lm,np.round(r,2),round(r.sum()*8,1))
print('--- NZ=12')
for NY in (8,9,10,11):
    for lm in ('pair','split'):
        r=simulate(NY,'zyx',lane_map=lm,NZ=12)
        print(NY,lm,np.round(r,2),round(r.sum()*8,1))
print('--- drift 0.15 cells (stale), NZ=8')
for NY in (9,10):
    for lm in ('pair','split'):
        r=simulate(NY,'zyx',lane_map=lm,drift=0.15)
        print(NY,lm,np.round(r,2),round(r.sum()*8,1))
