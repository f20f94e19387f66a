"""Short driver for ncu: two scheduling cycles of BASELINE config 2 (first = warm-up)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from yunikorn_k8shim_b200 import Engine, synth
masks = "--masks" in sys.argv
snap = synth.perf(masks=masks)
for rep in range(2):
    with Engine.for_snapshot(snap, batch=int(os.environ.get("YK_BATCH", "0"))) as e:
        ask, node, _ = e.cycle(snap.n_asks)
        print(rep, len(ask), e.stats()["sweep_ms"])
