import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yunikorn_k8shim_b200 import Engine, synth
from oracle import oracle_ctypes as oc

for snap in (synth.perf(), synth.perf(masks=True)):
    t = time.time(); want = oc.run(snap); t_or = time.time() - t
    for batch in (2048, 4096, 8192):
        for rep in range(2):
            with Engine.for_snapshot(snap, batch=batch) as e:
                t = time.time(); ask, node, _ = e.cycle(snap.n_asks); dt = time.time() - t
                st = e.stats()
            ok = np.array_equal(ask, want["ask"]) and np.array_equal(node, want["node"])
        print(f"{snap.name} batch={batch} ok={ok} cycle={dt*1e3:.1f}ms oracle={t_or*1e3:.1f}ms "
              f"sweep={st['sweep_ms']:.2f}ms sort={st['sort_ms']:.2f}ms commit={st['commit_ms']:.2f}ms "
              f"batches={st["batches"]} rows={st["rows_swept"]} evals/s(sweep)={st['evaluations']/max(st['sweep_ms'],1e-9)*1e3:.3e} "
              f"alloc/s={len(ask)/dt:.3e} host_ms={[round(x,2) for x in st['host_ms'][:8]]} dbg={st['dbg']} prof={[round(x/max(st['prof'][5],1)) for x in st['prof'][:5]]}")
