"""Randomized campaign over the engine's host side on CPU (no GPU): random cluster / queue-tree / gang shapes, batch
sizes, epoch lengths, row sharing, speculation on/off and max_bindings cuts through tests/host/engine_shim.cpp (real
orderer + ordered commit, CPU stand-in for the sweep), every result compared with the oracle.
Usage: python scripts/host_campaign.py <seed> <seconds>     (the CPU test suite runs a small fixed slice of this)"""
import os, subprocess
import sys, ctypes as C, time, random
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from yunikorn_k8shim_b200 import synth
from oracle import oracle_ctypes as oc
import test_engine_host as T
SO = "/tmp/yk_engine_shim_%d.so" % os.getpid()
subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", SO, os.path.join(ROOT, "tests", "host", "engine_shim.cpp")])
shim = C.CDLL(SO)
rng=random.Random(int(sys.argv[1]))
bad=0;n=0;t0=time.time()
while time.time()-t0 < float(sys.argv[2]):
    kind=rng.random()
    if kind<0.55:
        nn=rng.choice([2,3,4,6,9,15,40,120]); par=rng.randrange(1,5); lv=rng.randrange(1,4); apps=rng.randrange(1,4); tasks=rng.choice([3,8,20,60,100])
        s=synth.hier(nn,par,lv,apps,tasks,masks=rng.random()<0.3,priorities=rng.random()<0.6,seed=rng.randrange(10000),leaf_sort=rng.choice([1,0]),policy=rng.choice([0,1]),quota_frac=rng.choice([0.5,1.0,1.2,2.0,5.0]))
        if rng.random()<0.3:
            g=rng.choice([2,4,5])
            if tasks%g==0: s.ask_gang[:]=np.arange(s.n_asks)//g
        if rng.random()<0.4:
            s.q_prio_offset=np.array([rng.choice([0,0,5,-5,100]) for _ in range(s.n_queues)],dtype=np.int32)
            s.q_prio_fence=np.array([rng.random()<0.25 for _ in range(s.n_queues)],dtype=np.uint8)
    elif kind<0.75:
        s=synth.fuzz(rng.randrange(100000))
    elif kind<0.9:
        s=synth.poisoned_gangs(rng.randrange(10000),n_nodes=rng.choice([8,32,64]),n_gangs=rng.choice([10,40,80]),members=rng.choice([2,4,7]),policy=rng.choice([0,1]))
    else:
        s=synth.perf(rng.choice([3,10,50,300]),rng.randrange(1,12),rng.choice([5,40,120]),masks=rng.random()<0.5,policy=rng.choice([0,1]),seed=rng.randrange(10000))
        if rng.random()<0.3: s.ask_prio[:]=np.random.default_rng(rng.randrange(1000)).integers(-2,3,size=s.n_asks)
    if rng.random()<0.15:
        s=synth.redim(s,rng.choice([1,2,3,5,6,8]),rng.randrange(1000))
    try:
        want=oc.run(s)
    except RuntimeError:
        continue   # NaN score (zero total on a weighted dimension)
    gmax=np.bincount(s.ask_gang[s.ask_gang>=0]).max() if (s.ask_gang>=0).any() else 0
    for b in rng.sample([1,2,5,8,16,33,64,256,5000],4):
        if gmax>b: continue
        spec=rng.choice([0,1]); ep=rng.choice([None,None,1,7,50]); share=rng.choice([1,1,0,2])
        rc,ask,node,state,avail=T.run_engine_host(shim,s,batch=b,speculate=spec,epoch_limit=ep,share_rows=share)
        n+=1
        ok=rc==0 and np.array_equal(ask,want['ask']) and np.array_equal(node,want['node']) and np.array_equal(state,want['state']) and np.array_equal(avail,want['avail'])
        if not ok: bad+=1; print('FAIL',s.name,s.meta,b,spec,ep,share, flush=True)
    if rng.random()<0.4 and len(want['ask'])>2:
        k=rng.randrange(1,len(want['ask']))
        w2=oc.run(s,max_bindings=k)
        b=rng.choice([8,64,300])
        if gmax<=b:
            rc,ask,node,state,avail=T.run_engine_host(shim,s,batch=b,max_bindings=k,speculate=rng.choice([0,1]))
            n+=1
            if not (rc==0 and np.array_equal(ask,w2['ask']) and np.array_equal(node,w2['node']) and np.array_equal(avail,w2['avail'])): bad+=1; print('FAIL maxb',s.name,k,b,flush=True)
print('seed',sys.argv[1],'runs',n,'failures',bad,'%.0fs'%(time.time()-t0))
