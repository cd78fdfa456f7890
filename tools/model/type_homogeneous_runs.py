"""Block liveness and padding when the j stream is sorted by (xy row, type, z cell, Morton) as in MODE 1 of kernels_tile4.cuh.

Design-time numpy model (CPU only, a few minutes); not used by the product, the tests or the bench."""
import numpy as np, sys
rng=np.random.default_rng(1)
def run(N,W,R,radio=None,ratio=0.0,T=6,ntiles=30,name=''):
    a=np.ones(T) if radio is None else 1+np.array(radio)*ratio
    h=0.5*R*a
    Rmax=2*h.max()
    nc=int(W//(Rmax*(1+1e-5)))
    pos=rng.random((N,3))*W
    typ=rng.integers(0,T,N)
    c=np.minimum((pos*nc/W).astype(int),nc-1)
    sub=np.minimum((pos*4*nc/W).astype(int),4*nc-1)-4*c
    def spread(v): return (v&1)|((v&2)<<2)
    mort=spread(sub[:,2])|(spread(sub[:,1])<<1)|(spread(sub[:,0])<<2)
    cell=(c[:,0]*nc+c[:,1])*nc+c[:,2]
    key=cell*64+mort
    order=np.argsort(key,kind='stable')
    posA=pos[order];typA=typ[order];cellA=cell[order]
    startA=np.searchsorted(cellA,np.arange(nc**3+1))
    # B: key ((row*T+type)*nc+cz)*64+mort
    row=c[:,0]*nc+c[:,1]
    keyB=((row*T+typ)*nc+c[:,2])*64+mort
    oB=np.argsort(keyB,kind='stable')
    posB=pos[oB]; typB=typ[oB]; kB=(keyB[oB]//64)
    startB=np.searchsorted(kB,np.arange(nc*nc*T*nc+1))
    nblocks=0; live=0; acc=0; lanes=0; livehalf=0
    cells=rng.integers(0,nc**3,ntiles)
    for ce in cells:
        cz=ce%nc; cy=(ce//nc)%nc; cx=ce//(nc*nc)
        i0=startA[ce]; n=startA[ce+1]-i0
        if n==0: continue
        k=rng.integers(0,(n+127)//128)
        ib=i0+k*128; ni=min(n-k*128,128)
        for dx in (-1,0,1):
          for dy in (-1,0,1):
            x=(cx+dx)%nc; y=(cy+dy)%nc
            sx=(-W if cx+dx<0 else (W if cx+dx>=nc else 0)); sy=(-W if cy+dy<0 else (W if cy+dy>=nc else 0))
            segs=[(max(cz-1,0),min(cz+1,nc-1),0.0)]
            if cz==0: segs.append((nc-1,nc-1,-W))
            if cz==nc-1: segs.append((0,0,W))
            for z0,z1,sz in segs:
              for t in range(T):
                r=((x*nc+y)*T+t)*nc
                j0=startB[r+z0]; j1=startB[r+z1+1]
                if j1<=j0: continue
                pj=posB[j0:j1]+np.array([sx,sy,sz]); tj=typB[j0:j1]
                assert (tj==t).all()
                nj=j1-j0
                for l in range((ni+31)//32):
                    pi=posA[ib+32*l:ib+min(32*l+32,ni)]; ti=typA[ib+32*l:ib+min(32*l+32,ni)]
                    d=pj[None,:,:]-pi[:,None,:]
                    d2=(d*d).sum(-1)
                    cut=(h[ti][:,None]+h[tj][None,:])**2
                    ok=d2<cut
                    anyj=ok.any(0)
                    nq=(nj+3)//4
                    pad=np.zeros(nq*4,bool); pad[:nj]=anyj
                    nblocks+=nq; live+=pad.reshape(nq,4).any(1).sum(); acc+=ok.sum(); lanes+=len(pi)*nj
                    livehalf+=pad.reshape(nq*2,2).any(1).sum()
    print(name,'grid',nc,'per cell %.0f'%(N/nc**3),'quad blocks',nblocks,'live frac %.3f'%(live/nblocks),'half-live %.3f'%(livehalf/(2*nblocks)), 'force evals/accepted %.2f'%(live*128/acc), 'blocks per accepted-pair*128: %.2f'%(nblocks*128/acc), 'pad overhead %.3f'%(nblocks*128/lanes))
run(1000000,8000.,386.,name='pulser')
run(1000000,8000.,397.,radio=[1,.5,0,0,-.5,1],ratio=0.5,name='eater')
run(2000000,8000.,285.,T=8,name='settings2M')
