import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yunikorn_k8shim_b200 import Engine, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
s = {"cfg2": lambda: synth.perf(), "cfg3": lambda: synth.perf(masks=True), "ref": lambda: synth.reference_shape(), "cfg4": lambda: synth.hier()}[name]()
with Engine.for_snapshot(s, batch=4096, commit="host") as e:
    for it in range(3):
        if it:
            e.release(ask); e.load_snapshot(s); e.evaluate(0, 0)
        e.stats_reset()
        ask, node, _ = e.cycle(s.n_asks)
    st = e.stats()
    p = st["prof"]; n = max(p[5], 1)
    print(name, "commit_ms %.2f" % st["commit_ms"], "asks", p[5], "TSC cycles/ask: clean_scan %.0f walk %.0f choose+erase %.0f subtract+rescore %.0f reinsert %.0f" % tuple(x / n for x in p[:5]),
          "dbg words_scanned/ask %.2f walked/ask %.2f won_by_touched %.2f" % (st["dbg"][0] / n, st["dbg"][1] / n, st["dbg"][2] / n))
