#!/usr/bin/env python
"""How many Kuhn-Munkres roots of a BASELINE scene take the fast path (tightest column free, delta = 0) and how many need
the label / slack search?  CPU emulation on the oracle's cost matrices (no GPU): guides the voting-kernel work.
Result on cfg4 / cfg2 scenes (frame 6): every root takes the fast path -- the voting time of the positional-only
trackers is the one-warp serial walk over the rows, not the search."""
import sys, dataclasses, numpy as np
sys.path.insert(0, __file__.rsplit('/tools/', 1)[0])
import oracle
from similari_b200.workload import CONFIGS, Workload, tracker_options_for
def stats(name, frames=6):
    cfg=dataclasses.replace(CONFIGS[name], n_scenes=1)
    o=oracle.Tracker(tracker_options_for(name, oracle.make_options))
    wl=Workload(cfg)
    for fr in range(frames):
        f=wl.next_frame(); o.predict_batch(f["scene_ids"], f["det_offsets"], f["boxes"], features=f["features"])
    c=o.last_costs(int(f["scene_ids"][0]))
    m,n=c.shape
    thr=int(1.0*1e6) if name=="cfg4" else int(np.float32(0.3)*np.float32(1e6))
    # columns: first-seen order scanning entries by candidate then track
    order=[]; seen=set()
    for i in range(m):
        for j in range(n):
            if not np.isnan(c[i,j]) and j not in seen: seen.add(j); order.append(j)
    rest=[j for j in range(n) if j not in seen]
    cols=order+rest
    ny=m+n
    W=np.zeros((m,ny),dtype=np.int64)
    for i in range(m): W[i,i]=thr
    colpos={j:k for k,j in enumerate(cols)}
    ent=0
    for i in range(m):
        for j in range(n):
            if not np.isnan(c[i,j]): W[i,m+colpos[j]]=int(np.float32(c[i,j])*np.float32(1e6)); ent+=1
    lx=W.max(axis=1).astype(np.int64); ly=np.zeros(ny,np.int64)
    xy=-np.ones(m,int); yx=-np.ones(ny,int)
    fast=0; slow=0; iters=0; maxit=0
    for root in range(m):
        slack=lx[root]+ly-W[root]
        # fast path check
        y0=int(np.argmin(slack))
        if slack[y0]==0 and yx[y0]<0 and lx[root]>0:
            xy[root]=y0; yx[y0]=root; fast+=1; continue
        slow+=1
        alt=-np.ones(ny,int); S=np.zeros(m,bool); S[root]=True
        slackx=np.full(ny,root)
        it=0
        while True:
            it+=1
            free=alt<0
            sl=np.where(free,slack,np.iinfo(np.int64).max)
            y=int(np.argmin(sl)); delta=sl[y]; x=slackx[y]
            if delta>0:
                lx[S]-=delta; ly[~free]+=delta; slack[free]-=delta
            alt[y]=x
            if yx[y]<0:
                # augment
                while True:
                    xx=alt[y]; prec=xy[xx]; yx[y]=xx; xy[xx]=y; y=prec
                    if y<0: break
                break
            x2=yx[y]; S[x2]=True
            ns=lx[x2]+ly-W[x2]
            upd=(alt<0)&(ns<slack)
            slack=np.where(upd,ns,slack); slackx=np.where(upd,x2,slackx)
        iters+=it; maxit=max(maxit,it)
    print(name,"m",m,"n",n,"entries",ent,"entries/row",round(ent/m,2),"fast roots",fast,"slow roots",slow,"total search iterations",iters,"max",maxit)
stats("cfg4"); stats("cfg2")
