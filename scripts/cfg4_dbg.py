import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yunikorn_k8shim_b200 import Engine, synth
s = synth.hier()
with Engine.for_snapshot(s, batch=4096, commit="host") as e:
    for it in range(3):
        if it:
            e.release(ask); e.load_snapshot(s); e.evaluate(0, 0)
        e.stats_reset()
        t0 = time.perf_counter(); ask, node, _ = e.cycle(s.n_asks); dt = time.perf_counter() - t0
    st = e.stats()
    print(os.environ.get("YK_NO_EPOCH_ROWS"), "ms %.1f" % (dt * 1e3), {k: st[k] for k in ("batches", "nofit", "skipped", "sweep_launches", "rows_swept", "other_launches", "dbg", "d2h_bytes")}, [round(x, 2) for x in st["host_ms"]])
