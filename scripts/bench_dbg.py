"""debug: bench.py's value-mode loop with the per-phase host split, row sharing on / off"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from yunikorn_k8shim_b200 import Engine, synth

snap = synth.perf()
N, A, D = snap.n_nodes, snap.n_asks, snap.D
idxN, idxA = np.arange(N, dtype=np.uint32), np.arange(A, dtype=np.uint32)
totalT, availT = np.ascontiguousarray(snap.node_total.T), np.ascontiguousarray(snap.node_avail.T)
reqT = np.ascontiguousarray(snap.ask_req.T)
rank_arr = snap.node_rank()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
mode = sys.argv[1] if len(sys.argv) > 1 else "value"

for share in (True, False, True, False):
    eng = Engine(D=D, policy=snap.policy, weights=snap.weights, max_nodes=N, max_asks=A, max_apps=snap.n_apps,
                 max_queues=snap.n_queues, share_rows=share)
    eng.queues_set(snap.q_parent, snap.q_guaranteed, snap.q_max, snap.q_alloc, snap.q_sort)
    eng.apps_upsert(np.arange(snap.n_apps), snap.app_queue, snap.app_submit)
    prev = None
    for step in range(8):
        if prev is not None:
            eng.release(prev)
        if mode == "e2e":
            flush.add_(1); torch.cuda.synchronize()
        t_u = time.perf_counter()
        eng.nodes_upsert(idxN, totalT, availT, snap.node_taint, snap.node_label, rank_arr, snap.node_flags)
        eng.asks_upsert(idxA, reqT, snap.ask_app, snap.ask_create, snap.ask_tol, snap.ask_need, snap.ask_deny,
                        snap.ask_prio, snap.ask_node, snap.ask_flags, snap.ask_gang)
        t_u = time.perf_counter() - t_u
        if mode == "value":
            eng.evaluate(0, 0)
            flush.add_(1); torch.cuda.synchronize()
        eng.stats_reset()
        t0 = time.perf_counter()
        ask, node, _ = eng.cycle(A)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = eng.stats()
        prev = ask
        if step >= 3:
            print(f"{mode} share={share} step={step} cycle={dt*1e3:.2f}ms total_ms={st["total_ms"]:.2f} upsert={t_u*1e3:.2f}ms commit={st['commit_ms']:.2f} rows={st['rows_swept']} "
                  f"host_ms={[round(x, 2) for x in st['host_ms'][:8]]}")
    eng.close()
