"""A few scheduling cycles of a bench shape under ncu (profiles/): env YK_NO_ROW_SHARING / YK_BATCH as for the engine.
usage: prof_cycle.py [config2|config3|reference] [host|device]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yunikorn_k8shim_b200 import Engine, synth

name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else ("config3" if "--masks" in sys.argv else "config2")
commit = sys.argv[2] if len(sys.argv) > 2 else "host"
s = {"config2": lambda: synth.perf(), "config3": lambda: synth.perf(masks=True), "reference": lambda: synth.reference_shape()}[name]()
with Engine.for_snapshot(s, batch=int(os.environ.get("YK_BATCH", "0")), commit=commit) as e:
    for _ in range(2):
        ask, node, _ = e.cycle(s.n_asks)
        e.release(ask)
        e.load_snapshot(s)
    st = e.stats()
    print(len(ask), st["sweep_launches"], st["rows_swept"], st["lattice_launches"])
