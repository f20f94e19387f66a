"""Quick timing of one cycle per commit mode on the bench shapes (development aid; bench.py is the contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yunikorn_k8shim_b200 import Engine, synth
from oracle import oracle_ctypes as oc

which = sys.argv[1:] or ["cfg2", "cfg3", "ref", "cfg5", "cfg4"]
shapes = {"cfg2": lambda: synth.perf(), "cfg3": lambda: synth.perf(masks=True), "ref": lambda: synth.reference_shape(),
          "cfg5": lambda: synth.gangs(), "cfg4": lambda: synth.hier()}
for name in which:
    s = shapes[name]()
    want = oc.run(s)
    for commit in ("device", "host"):
        with Engine.for_snapshot(s, batch=4096, commit=commit) as e:
            best = 1e9
            for it in range(4):
                if it:
                    e.release(ask)
                    e.load_snapshot(s)
                    e.evaluate(0, 0)
                e.stats_reset()
                t0 = time.perf_counter()
                ask, node, _ = e.cycle(s.n_asks)
                dt = time.perf_counter() - t0
                best = min(best, dt)
            st = e.stats()
        ok = np.array_equal(ask, want["ask"]) and np.array_equal(node, want["node"])
        print(f"{name} {commit}: {best*1e3:.2f} ms/cycle  {len(ask)/best/1e6:.2f} M alloc/s identical={ok} "
              f"lattice: launches={st['lattice_launches']} subruns={st['lattice_subruns']} asks={st['lattice_asks']} "
              f"elems={st['lattice_elements']} sorts={st['lattice_sorts']} full={st['lattice_fullscans']} handoffs={st['lattice_handoffs']} "
              f"ms={st['lattice_ms']:.2f} uniform: runs={st['uniform_runs']} asks={st['uniform_asks']} retries={st['uniform_retries']} host_ms={[round(x, 2) for x in st['host_ms']]} total_ms={st['total_ms']:.2f} | sweeps={st['sweep_launches']} commit_ms={st['commit_ms']:.2f} d2h={st['d2h_bytes']}", flush=True)
