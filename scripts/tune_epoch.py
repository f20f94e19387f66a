"""cycle time of config 2 as a function of sweep batch size and epoch length (YK_EPOCH_NODES), both policies"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yunikorn_k8shim_b200 import Engine, synth

for policy in (synth.POLICY_FAIR, synth.POLICY_BINPACKING):
    snap = synth.perf(policy=policy)
    ref = None
    for batch in (2048, 4096, 8192, 16384):
        for epoch in (0, 4096, 16384, 65536):
            if epoch:
                os.environ["YK_EPOCH_NODES"] = str(epoch)
            else:
                os.environ.pop("YK_EPOCH_NODES", None)
            best, cm = None, None
            for rep in range(4):
                with Engine.for_snapshot(snap, batch=batch) as e:
                    t = time.perf_counter(); ask, node, _ = e.cycle(snap.n_asks); dt = time.perf_counter() - t
                    st = e.stats()
                if best is None or dt < best:
                    best, cm = dt, st["commit_ms"]
            if ref is None:
                ref = (ask.copy(), node.copy())
            same = np.array_equal(ask, ref[0]) and np.array_equal(node, ref[1])
            print(f"policy={policy} batch={batch} epoch={epoch or 'default'} cycle={best*1e3:.2f}ms commit={cm:.2f}ms "
                  f"batches={st['batches']} same={same}", flush=True)
