"""cycle time of config 2 / 3 under both node-sort policies (fair, binpacking), checked against the oracle"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yunikorn_k8shim_b200 import Engine, synth
from oracle import oracle_ctypes as oc

for masks in (False, True):
    for policy in (synth.POLICY_FAIR, synth.POLICY_BINPACKING):
        snap = synth.perf(masks=masks, policy=policy)
        t = time.time(); want = oc.run(snap); t_or = time.time() - t
        best = None
        for rep in range(3):
            with Engine.for_snapshot(snap, batch=4096) as e:
                t = time.time(); ask, node, _ = e.cycle(snap.n_asks); dt = time.time() - t
                st = e.stats()
            best = dt if best is None else min(best, dt)
        ok = np.array_equal(ask, want["ask"]) and np.array_equal(node, want["node"])
        print(f"{snap.name} policy={'binpacking' if policy else 'fair'} ok={ok} cycle={best*1e3:.1f}ms oracle={t_or*1e3:.0f}ms "
              f"commit={st['commit_ms']:.2f}ms rows={st['rows_swept']} walk_examined={st['dbg'][1]} alloc/s={len(ask)/best:.3e}")
