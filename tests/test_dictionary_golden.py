"""The reference's own known-answer tables at the PredicateManager boundary
(/root/reference/pkg/plugin/predicates/predicate_manager_test.go TestPodFitsSelector, TestReserveNodeSelector,
TestReserveAlloc), pushed through the product's dictionary encoder (strings -> bit sets, csrc/yk_dict.cpp) and then
through (a) the CPU oracle's predicate and (b) on the GPU, the device predicate (yk_evaluate)."""
import json
import os

import numpy as np
import pytest

from yunikorn_k8shim_b200 import synth
from yunikorn_k8shim_b200.dictionary import Dictionary

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pod_fits_selector.json")))


def encode_selector_case(c):
    d = Dictionary()
    name = c.get("nodeName") or "machine1"
    pod = c["pod"]
    aff = pod.get("affinity", "absent")
    if aff == "absent" or aff is None:
        m = d.pod(node_selector=pod.get("nodeSelector"))
    else:
        m = d.pod(node_selector=pod.get("nodeSelector"), affinity_terms=aff["terms"], has_affinity=True)
    # the node is registered AFTER the pod allocated its bits (and a decoy before, to exercise re-evaluation)
    d.node(1, "decoy", {"unrelated": "x"})
    lb, tb = d.node(0, name, c["labels"])
    return m, lb, tb, d


def encode_taint_case(c):
    d = Dictionary()
    lb, tb = d.node(0, "foo", {}, c["taints"], c["unschedulable"])
    m = d.pod(tolerations=c["tolerations"])
    return m, lb, tb, d


def as_snapshot(m, lb, tb):
    D = 3
    total = np.array([[10, 20, 32], [10, 20, 32]], dtype=np.int64)
    z = np.zeros(1, dtype=np.uint64)
    node = -1 if m.required_node == 0xFFFFFFFF else int(m.required_node)
    return synth._finish("golden", D, 0, total, total.copy(), np.array([tb, 0], np.uint64), np.array([lb, 0], np.uint64),
                         ["n0", "n1"], synth._single_queue(D), np.ones(1, np.int32), np.zeros(1, np.int32),
                         np.array([[0, 0, 1]], dtype=np.int64), z + np.uint64(m.tolerated_bits), z + np.uint64(m.required_bits),
                         z + np.uint64(m.forbidden_bits), ask_node=np.array([node], np.int32))


@pytest.mark.parametrize("c", GOLD["selector_cases"], ids=[c["name"] for c in GOLD["selector_cases"]])
def test_selector_table_encoder_plus_oracle(oracle, c):
    m, lb, tb, _ = encode_selector_case(c)
    assert m.flags == 0
    rc = oracle.predicate(as_snapshot(m, lb, tb), 0, 0)
    assert (rc == 0) == c["fits"], (c["name"], rc, hex(lb), hex(m.required_bits), hex(m.forbidden_bits))
    if not c["fits"]:
        assert rc == 7   # NodeAffinity


@pytest.mark.parametrize("c", GOLD["taint_cases"], ids=[c["name"][:40] for c in GOLD["taint_cases"]])
def test_taint_table_encoder_plus_oracle(oracle, c):
    m, lb, tb, _ = encode_taint_case(c)
    rc = oracle.predicate(as_snapshot(m, lb, tb), 0, 0)
    assert (rc == 0) == c["fits"], (c["name"], rc)
    if not c["fits"]:
        assert rc == 6   # TaintToleration / NodeUnschedulable


def test_encoder_semantics_beyond_the_table():
    d = Dictionary()
    d.node(0, "a", {"zone": "z1", "n": "5"}, [{"key": "k", "value": "v", "effect": "NoExecute"},
                                                {"key": "soft", "value": "", "effect": "PreferNoSchedule"}])
    d.node(1, "b", {"zone": "z2"})
    lb0, tb0 = d.node_bits(0)
    assert bin(tb0).count("1") == 1                           # PreferNoSchedule is not a filter
    g = d.generation
    m = d.pod(affinity_terms=[{"expr": [{"key": "n", "op": "Lt", "values": ["7"]}, {"key": "zone", "op": "NotIn", "values": ["z2"]}]}])
    assert d.generation > g                                   # new bits -> nodes must be re-read
    lb0, _ = d.node_bits(0)
    lb1, _ = d.node_bits(1)
    fit = lambda lb: (lb & m.required_bits) == m.required_bits and (lb & m.forbidden_bits) == 0  # noqa: E731
    assert fit(lb0) and not fit(lb1)
    # toleration matching: effect-specific, key-specific, Equal vs Exists, tolerate-everything
    t = lambda **kw: d.pod(tolerations=[kw]).tolerated_bits   # noqa: E731
    assert t(key="k", op="Equal", value="v", effect="NoExecute") & tb0 == tb0
    assert t(key="k", op="Equal", value="other", effect="NoExecute") & tb0 == 0
    assert t(key="k", op="Exists", effect="NoSchedule") & tb0 == 0       # wrong effect
    assert t(key="k", op="Exists") & tb0 == tb0                          # empty effect = all effects
    assert t(op="Exists") == 0xFFFFFFFFFFFFFFFF                          # empty key + Exists tolerates everything, also future taints
    # pod.Spec.NodeName: known node -> index, unknown node -> impossible bit
    assert d.pod(node_name="b").required_node == 1
    assert d.pod(node_name="nope").required_bits >> 63 == 1
    # dictionary exhaustion is flagged, never approximated
    for i in range(70):
        last = d.pod(node_selector={f"k{i}": "v"})
    assert last.flags & 1


@pytest.mark.gpu
def test_selector_and_taint_tables_on_device():
    """same golden tables, answered by the device predicate through the C ABI (yk_evaluate)"""
    from yunikorn_k8shim_b200 import Engine
    cases = [(c, encode_selector_case(c)) for c in GOLD["selector_cases"]] + [(c, encode_taint_case(c)) for c in GOLD["taint_cases"]]
    for c, (m, lb, tb, _) in cases:
        s = as_snapshot(m, lb, tb)
        with Engine.for_snapshot(s) as e:
            rc = e.evaluate(0, 0)
            assert (rc == 0) == c["fits"], (c["name"], rc)
            ask, node, _ = e.cycle(1)                          # and the sweep + commit agree: bound to node 0 or not at all
            assert (len(ask) == 1 and node[0] == 0) == c["fits"] or (len(ask) == 1 and node[0] == 1 and not c["fits"]), c["name"]
