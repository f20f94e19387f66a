"""One device-commit cycle of a bench shape (for ncu captures of yk_lattice_kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yunikorn_k8shim_b200 import Engine, synth
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
s = {"cfg2": lambda: synth.perf(), "cfg3": lambda: synth.perf(masks=True), "ref": lambda: synth.reference_shape()}[name]()
with Engine.for_snapshot(s, batch=4096, commit="device") as e:
    ask, node, _ = e.cycle(s.n_asks)
    print(len(ask), e.stats()["lattice_ms"])
