import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yunikorn_k8shim_b200 import Engine, synth
from oracle import oracle_ctypes as oc
bad = 0
for batch in (7, 64, 1024):
    for seed in range(60):
        s = synth.fuzz(seed); want = oc.run(s)
        try:
            with Engine.for_snapshot(s, batch=batch) as e:
                ask, node, _ = e.cycle(s.n_asks); st = e.ask_states(np.arange(s.n_asks))
        except Exception as exc:
            print('seed', seed, 'batch', batch, 'EXC', exc); continue
        ok = np.array_equal(ask, want['ask']) and np.array_equal(node, want['node']) and np.array_equal(st, want['state'])
        if not ok:
            bad += 1
            k = next((i for i in range(min(len(ask), len(want['ask']))) if ask[i] != want['ask'][i] or node[i] != want['node'][i]), None)
            d = np.nonzero(st != want['state'])[0]
            print('seed', seed, 'batch', batch, 'len', len(ask), len(want['ask']), 'first diff at', k,
                  'got', (ask[k], node[k]) if k is not None and k < len(ask) else None, 'want', (want['ask'][k], want['node'][k]) if k is not None else None,
                  'state diffs', d[:6], st[d[:6]], want['state'][d[:6]], 'gangs', int((s.ask_gang>=0).any()), 'policy', s.policy, flush=True)
print('mismatches', bad, 'env', os.environ.get('YK_NO_SPECULATION'), os.environ.get('YK_EPOCH_NODES'))
