"""Randomized cross-check of the label / taint dictionary encoder (csrc/yk_dict.cpp, include/ykgpu_dict.h): for random
nodes and pods built from a small vocabulary, the verdict computed from the 64-bit masks
    (taint & ~tolerated) == 0 and (label & required) == required and (label & forbidden) == 0 and node-name match
must equal the verdict of the string-level restatement of the Kubernetes plugins (oracle/py_k8s_predicates.py) for every
(pod, node) pair.  The oracle is first checked against the reference's own selector table.  Host code: no GPU needed."""
import json
import os
import random

import pytest

from oracle import py_k8s_predicates as k8s
from yunikorn_k8shim_b200.dictionary import Dictionary

HERE = os.path.dirname(os.path.abspath(__file__))
NONE = 0xFFFFFFFF
SLOWPATH = 1

KEYS = ["zone", "arch", "disk", "gpu", "cores", "tier"]
VALUES = {"zone": ["a", "b", "c"], "arch": ["amd64", "arm64"], "disk": ["ssd", "hdd"], "gpu": ["a100", "b200"],
          "cores": ["8", "16", "64", "many"], "tier": ["1", "2", "x"]}
TAINT_KEYS = ["dedicated", "gpu-only", "maintenance"]
EFFECTS = ["NoSchedule", "NoExecute", "PreferNoSchedule"]


def test_oracle_agrees_with_the_reference_selector_table():
    gold = json.load(open(os.path.join(HERE, "golden", "pod_fits_selector.json")))
    n = 0
    for c in gold["selector_cases"]:
        aff = c["pod"].get("affinity", "absent")
        pod = {"node_selector": c["pod"].get("nodeSelector")}
        if aff != "absent" and aff is not None:
            pod["has_affinity"], pod["affinity_terms"] = True, aff["terms"]
        got = k8s.node_affinity(pod, c.get("labels") or {}, c.get("nodeName") or "machine1")
        assert got == c["fits"], c["name"]
        n += 1
    for c in gold["taint_cases"]:
        assert k8s.taint_toleration({"tolerations": c["tolerations"]}, c["taints"], c["unschedulable"]) == c["fits"], c["name"]
    assert n >= 27


def _rand_node(rng, i):
    labels = {k: rng.choice(VALUES[k]) for k in KEYS if rng.random() < 0.6}
    taints = []
    for k in TAINT_KEYS:
        if rng.random() < 0.25:
            taints.append({"key": k, "value": rng.choice(["", "x", "y"]), "effect": rng.choice(EFFECTS)})
    return {"name": f"node-{i}", "labels": labels, "taints": taints, "unschedulable": rng.random() < 0.1}


def _rand_req(rng):
    k = rng.choice(KEYS)
    op = rng.choice(["In", "In", "NotIn", "Exists", "DoesNotExist", "Gt", "Lt"])
    if op in ("In", "NotIn"):
        return {"key": k, "op": op, "values": rng.sample(VALUES[k], rng.randrange(1, len(VALUES[k]) + 1))}
    if op in ("Gt", "Lt"):
        return {"key": rng.choice(["cores", "tier"]), "op": op, "values": [str(rng.choice([0, 1, 8, 16, 32, 100]))]}
    return {"key": k, "op": op, "values": []}


def _rand_pod(rng, n_nodes):
    pod = {}
    if rng.random() < 0.4:
        pod["node_selector"] = {k: rng.choice(VALUES[k]) for k in rng.sample(KEYS, rng.randrange(1, 3))}
    if rng.random() < 0.6:
        pod["has_affinity"] = True
        terms = []
        for _ in range(rng.randrange(0, 4)):
            t = {"expr": [_rand_req(rng) for _ in range(rng.randrange(0, 3))]}
            if rng.random() < 0.2:
                t["fields"] = [{"key": "metadata.name", "op": rng.choice(["In", "NotIn"]), "values": [f"node-{rng.randrange(n_nodes + 1)}"]}]
            terms.append(t)
        pod["affinity_terms"] = terms
    tols = []
    for k in TAINT_KEYS + ["", k8s.UNSCHEDULABLE_KEY]:
        if rng.random() < 0.3:
            op = rng.choice(["", "Equal", "Exists"]) if k else "Exists"
            tols.append({"key": k, "op": op, "value": rng.choice(["", "x", "y"]) if op != "Exists" else "",
                         "effect": rng.choice(["", "NoSchedule", "NoExecute", "PreferNoSchedule"])})
    pod["tolerations"] = tols
    if rng.random() < 0.1:
        pod["node_name"] = f"node-{rng.randrange(n_nodes + 1)}"          # may name a node that does not exist
    return pod


@pytest.mark.parametrize("seed", range(40))
def test_masks_give_the_string_level_verdict(seed):
    rng = random.Random(seed)
    n_nodes = rng.randrange(3, 14)
    nodes = [_rand_node(rng, i) for i in range(n_nodes)]
    pods = [_rand_pod(rng, n_nodes) for _ in range(rng.randrange(5, 30))]
    d = Dictionary()
    try:
        for i, nd in enumerate(nodes):
            d.node(i, nd["name"], nd["labels"], nd["taints"], nd["unschedulable"])
        masks = [d.pod(node_selector=p.get("node_selector"), affinity_terms=p.get("affinity_terms"),
                       has_affinity=p.get("has_affinity", False), tolerations=p.get("tolerations"), node_name=p.get("node_name"))
                 for p in pods]
        bits = [d.node_bits(i) for i in range(n_nodes)]               # after all pods: every bit they allocated is known
        checked = 0
        for p, m in zip(pods, masks):
            if m.flags & SLOWPATH:
                continue                                               # dictionary full: the pod goes to the Go path
            for i, nd in enumerate(nodes):
                lb, tb = bits[i]
                got = (tb & ~m.tolerated_bits) == 0 and (lb & m.required_bits) == m.required_bits and \
                    (lb & m.forbidden_bits) == 0 and (m.required_node == NONE or m.required_node == i)
                want = k8s.fits(p, nd)
                assert got == want, (seed, p, nd, hex(lb), hex(tb), hex(m.tolerated_bits), hex(m.required_bits), hex(m.forbidden_bits), m.required_node)
                checked += 1
        assert checked > 0
    finally:
        d.close()


def test_full_dictionary_flags_slow_path_and_stays_exact():
    """a vocabulary far larger than 63 requirement bits: pods that no longer get bits are flagged YK_ASK_SLOWPATH (never
    approximated), and every pod that did get its bits still yields the string-level verdict"""
    rng = random.Random(3)
    keys = [f"k{i}" for i in range(14)]
    saved = (list(KEYS), dict(VALUES))
    try:
        KEYS[:] = keys
        VALUES.clear()
        VALUES.update({k: [f"v{j}" for j in range(6)] for k in keys})
        VALUES["cores"], VALUES["tier"] = ["8", "16"], ["1", "2"]
        nodes = [_rand_node(rng, i) for i in range(10)]
        pods = [_rand_pod(rng, 10) for _ in range(400)]
    finally:
        KEYS[:] = saved[0]
        VALUES.clear()
        VALUES.update(saved[1])
    d = Dictionary()
    try:
        for i, nd in enumerate(nodes):
            d.node(i, nd["name"], nd["labels"], nd["taints"], nd["unschedulable"])
        masks = [d.pod(node_selector=p.get("node_selector"), affinity_terms=p.get("affinity_terms"),
                       has_affinity=p.get("has_affinity", False), tolerations=p.get("tolerations"), node_name=p.get("node_name"))
                 for p in pods]
        bits = [d.node_bits(i) for i in range(10)]
        slow = sum(1 for m in masks if m.flags & SLOWPATH)
        assert 50 < slow < 390
        for p, m in zip(pods, masks):
            if m.flags & SLOWPATH:
                continue
            for i, nd in enumerate(nodes):
                lb, tb = bits[i]
                got = (tb & ~m.tolerated_bits) == 0 and (lb & m.required_bits) == m.required_bits and \
                    (lb & m.forbidden_bits) == 0 and (m.required_node == NONE or m.required_node == i)
                assert got == k8s.fits(p, nd), (p, nd)
    finally:
        d.close()
