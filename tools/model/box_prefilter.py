"""How many (layer, j-group) combinations survive a bounding-box test, for j groups of 4 / 8 / 16 / 32: the numbers behind the quad-level box prefilter of kernels_tile4.cuh.

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
    pos=pos[order];typ=typ[order];cell=cell[order]
    start=np.searchsorted(cell,np.arange(nc**3+1))
    G=[4,8,16,32]
    tot={g:dict(groups=0,boxlive=0,quads_in_live=0, lanelive=0) for g in G}
    nblocks=0; live=0; acc=0
    cells=rng.integers(0,nc**3,ntiles)
    for ce in cells:
        cz=ce%nc; cy=(ce//nc)%nc; cx=ce//(nc*nc)
        i0=start[ce]; n=start[ce+1]-i0
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
                row=(x*nc+y)*nc
                j0=start[row+z0]; j1=start[row+z1+1]
                if j1<=j0: continue
                pj=pos[j0:j1]+np.array([sx,sy,sz]); tj=typ[j0:j1]
                nj=j1-j0
                for l in range((ni+31)//32):
                    pi=pos[ib+32*l:ib+min(32*l+32,ni)]; ti=typ[ib+32*l:ib+min(32*l+32,ni)]
                    lo_i=pi.min(0); hi_i=pi.max(0); Hi=h[ti].max()
                    d=pj[None,:,:]-pi[:,None,:]
                    d2=(d*d).sum(-1)
                    cut=(h[ti][:,None]+h[tj][None,:])**2
                    ok=d2<cut
                    anyj=ok.any(0)
                    nq=(nj+3)//4
                    pad=np.zeros(nq*4,bool); pad[:nj]=anyj
                    nblocks+=nq; live+=pad.reshape(nq,4).any(1).sum(); acc+=ok.sum()
                    for g in G:
                        ng=(nj+g-1)//g
                        for q in range(ng):
                            s=slice(q*g,min(q*g+g,nj))
                            lo_j=pj[s].min(0); hi_j=pj[s].max(0); Hj=h[tj[s]].max()
                            gap=np.maximum(0,np.maximum(lo_j-hi_i,lo_i-hi_j))
                            bl=(gap*gap).sum()<(Hi+Hj)**2
                            tot[g]['groups']+=1
                            if bl:
                                tot[g]['boxlive']+=1
                                tot[g]['quads_in_live']+=( (min(q*g+g,nj)-q*g)+3)//4
    print(name,'grid',nc,'per cell',N/nc**3,'quad blocks',nblocks,'live frac %.3f'%(live/nblocks))
    for g in G:
        t=tot[g]
        print('  group %2d: box-live frac %.3f ; quad-blocks needing fine test %.3f of all'%(g,t['boxlive']/t['groups'],t['quads_in_live']/nblocks))
run(1000000,8000.,386.,name='pulser')
run(1000000,8000.,397.,radio=[1,.5,0,0,-.5,1],ratio=0.5,name='eater')
run(2000000,8000.,285.,T=8,name='settings2M')
