"""Regenerates the oracle-produced golden bindings under tests/golden/.

These are NOT reference outputs: the Go reference cannot run in this image and no reference test pins a
pod->node map (SURVEY.md 8c "parity unpinned"), so the goldens freeze what the C++ oracle (cross-checked
against oracle/py_oracle.py) produces for the seeded BASELINE config-1 snapshots.  They guard against
accidental drift of the oracle itself.  Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle_ctypes as oc, py_oracle  # noqa: E402
from yunikorn_k8shim_b200 import synth  # noqa: E402

out = {"_generated_by": "tests/golden/make_golden.py", "_status": "oracle-generated (parity unpinned vs the Go reference)", "cases": []}
for variant in ("bare", "sized"):
    for policy in (synth.POLICY_FAIR, synth.POLICY_BINPACKING):
        s = synth.kwok(100, 10, 50, variant=variant, policy=policy)
        r = oc.run(s)
        p = py_oracle.run(s)
        assert list(r["ask"]) == p["ask"] and list(r["node"]) == p["node"]
        out["cases"].append({"generator": "kwok", "args": {"n_nodes": 100, "n_apps": 10, "replicas": 50, "variant": variant, "policy": policy},
                             "n_bindings": len(r["ask"]), "hash": f"{oc.bindings_hash(r['ask'], r['node']):#x}",
                             "node_id_of_binding": [s.node_id[n] for n in r["node"]],
                             "ask": [int(a) for a in r["ask"]]})
for kw in ({"n_nodes": 10000, "n_apps": 400, "tasks": 125, "masks": False}, {"n_nodes": 10000, "n_apps": 400, "tasks": 125, "masks": True}):
    s = synth.perf(**kw)
    r = oc.run(s)
    out["cases"].append({"generator": "perf", "args": kw, "n_bindings": len(r["ask"]),
                         "hash": f"{oc.bindings_hash(r['ask'], r['node']):#x}"})
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "oracle_bindings.json"), "w"), indent=0)
print("wrote", len(out["cases"]), "cases")
