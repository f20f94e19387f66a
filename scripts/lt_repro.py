import sys, os, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yunikorn_k8shim_b200 import Engine, synth
from oracle import oracle_ctypes as oracle
s = synth.perf(80, 4, 60, masks=True, seed=51)
s.ask_tol[:] = s.ask_tol[0]; s.ask_need[:] = 0; s.ask_deny[:] = 0
first = oracle.run(s, max_bindings=50)
for commit in sys.argv[1:] or ["device"]:
    with Engine.for_snapshot(s, batch=32, commit=commit) as e:
        ask, node, _ = e.cycle(50)
        print(commit, "cycle1", np.array_equal(ask, first["ask"]) and np.array_equal(node, first["node"]), flush=True)
        t = copy.deepcopy(s)
        t.node_avail = e.nodes_available(np.arange(s.n_nodes))
        t.ask_flags[first["ask"]] = 1
        pend = np.setdiff1d(np.arange(s.n_asks), first["ask"])
        chg = pend[::3]
        t.ask_tol[chg] = np.uint64(0xFFFF)
        t.ask_need[chg] = np.uint64(1) << (np.arange(len(chg)) % 8).astype(np.uint64)
        t.ask_req[chg[::2], 0] += 30
        e.asks_upsert(chg, t.ask_req[chg], t.ask_app[chg], t.ask_create[chg], t.ask_tol[chg], t.ask_need[chg], t.ask_deny[chg])
        want = oracle.run(t)
        ask2, node2, _ = e.cycle(s.n_asks)
        print(commit, "cycle2", np.array_equal(ask2, want["ask"]) and np.array_equal(node2, want["node"]), e.stats()["lattice_fullscans"], flush=True)
