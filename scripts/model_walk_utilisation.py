#!/usr/bin/env python
"""CPU model of the small-quad walk's lane utilisation (no GPU needed): projects every patch of the built-in scene
through a shooter's five face matrices as the set-up kernel does, takes the small-quad records in (approximate) queue
order and reports, per shooter, how full the quarter-warp walk of raster_queue_kernel runs:
  quarter balance   steps a quarter warp needs / steps of the longest of the warp's four records
  last-step         positions / (8 x steps)
  coverage          covered pixels / visited pixels (two per position)
    python scripts/model_walk_utilisation.py [area] [hemicube]     (defaults: 0.014 512 = BASELINE config 2)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc
from radiosity_b200 import api
def face_win(f,N):
    if f==0: return 0,N, 0,N,N,N//2
    if f==1: return N,N//2, N,N,N,N//2
    if f==2: return -(N//2),0, 0,0,N//2,N
    if f==3: return N+N//2,0, N+N//2,0,N//2,N
    return N//2,0, N//2,0,N,N
def records(v, sh, N):
    """approximate small-quad records (bw,bh,covered) in queue order for shooter sh"""
    q=v.reshape(-1,4,3).astype(np.float64)
    P=q.shape[0]
    out=[]
    for f in range(5):
        m=orc.mvp(v[sh], api.FACE_TO_LOOK[f]).astype(np.float64).reshape(4,4).T
        vpx,vpy,scx,scy,scw,sch=face_win(f,N)
        hp=np.concatenate([q,np.ones((P,4,1))],-1)@m.T   # P,4,4
        w=hp[...,3]
        ok=((hp[...,2]+w)>=0).all(1)
        with np.errstate(all='ignore'):
            X=np.rint((hp[...,0]/w*N/2+vpx+N/2)*256); Y=np.rint((hp[...,1]/w*N/2+vpy+N/2)*256)
        def area(i,j,k): return (X[:,j]-X[:,i])*(Y[:,k]-Y[:,i])-(Y[:,j]-Y[:,i])*(X[:,k]-X[:,i])
        aA=area(0,1,2); aB=area(0,2,3)
        front=ok&(aA>0)&(aB>0)
        minx=X.min(1); maxx=X.max(1); miny=Y.min(1); maxy=Y.max(1)
        px0=np.maximum(np.ceil((minx-128)/256),scx); px1=np.minimum(np.floor((maxx-128)/256),scx+scw-1)
        py0=np.maximum(np.ceil((miny-128)/256),scy); py1=np.minimum(np.floor((maxy-128)/256),scy+sch-1)
        vis=front&(px0<=px1)&(py0<=py1)
        bw=(px1-px0+1); bh=(py1-py0+1)
        cov=(aA+aB)/2/65536
        idx=np.nonzero(vis)[0]
        for i in idx: out.append((i//32, f, i, int(bw[i]), int(bh[i]), float(cov[i])))
    out.sort()
    return np.array([(o[3],o[4],o[5]) for o in out])
if __name__=="__main__":
    area=float(sys.argv[1]) if len(sys.argv)>1 else 0.014
    N=int(sys.argv[2]) if len(sys.argv)>2 else 512
    v,c,r,il=orc.scene_cornell(area)
    rng=np.random.default_rng(3)
    for sh in rng.choice(v.shape[0],5,replace=False):
        R=records(v,int(sh),N)
        bw,bh,cov=R[:,0],R[:,1],R[:,2]
        small=(bw*bh<=512)&(bw<=255)&(bh<=255)&(bw*bh>2)
        bw,bh,cov=bw[small],bh[small],cov[small]
        npx=bw*np.ceil(bh/2); steps=np.ceil(npx/8)
        n4=len(steps)//4*4
        g=steps[:n4].reshape(-1,4)
        U1=g.sum()/(4*g.max(1).sum())
        ss=np.sort(steps[:n4]).reshape(-1,4); U1s=ss.sum()/(4*ss.max(1).sum())
        U2=npx.sum()/(8*steps.sum())
        U3=np.minimum(cov,bw*bh).sum()/(2*npx.sum())
        print(f"shooter {int(sh)}: records {len(R)} small {small.sum()} mean bbox {np.mean(bw*bh):.0f}px steps mean {steps.mean():.1f} | quarter balance {U1:.2f} (sorted {U1s:.2f}) last-step {U2:.2f} coverage {U3:.2f} -> total {U1*U2*U3:.2f}")
