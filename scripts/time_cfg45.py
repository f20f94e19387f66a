import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yunikorn_k8shim_b200 import Engine, synth
from oracle import oracle_ctypes as oc
for snap in (synth.hier(), synth.gangs()):
    t = time.time(); want = oc.run(snap); t_or = time.time() - t
    for batch in (2048, 4096, 8192):
        with Engine.for_snapshot(snap, batch=batch) as e:
            e.evaluate(0, 0)
            t = time.time(); ask, node, _ = e.cycle(snap.n_asks); dt = time.time() - t
            st = e.stats()
        ok = np.array_equal(ask, want["ask"]) and np.array_equal(node, want["node"])
        print(f"{snap.name} batch={batch} ok={ok} n={len(ask)} cycle={dt*1e3:.1f}ms oracle={t_or*1e3:.1f}ms sweep={st['sweep_ms']:.2f}ms "
              f"commit={st['commit_ms']:.2f}ms batches={st['batches']} evals={st['evaluations']:.3e} "
              f"host_ms={[round(x,1) for x in st['host_ms'][:8]]} dbg={st['dbg']} d2h={st['d2h_bytes']/1e6:.0f}MB", flush=True)
