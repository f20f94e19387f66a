"""Pins the CPU oracle: (1) the reference's own known-answer tables at the PredicateManager boundary,
(2) hand-computed vectors for the [EXT] core arithmetic (node score, DRF share comparison, NodeID order),
(3) agreement with the independent pure-Python restatement, (4) frozen oracle-generated goldens."""
import json
import os

import numpy as np
import pytest

from yunikorn_k8shim_b200 import synth
from oracle import py_oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def one_pair_snapshot(alloc, used, req, pod_node, node_name):
    """a 2-node snapshot (the named node + a decoy) with one ask"""
    D = 3
    total = np.array([alloc, alloc], dtype=np.int64)
    avail = total - np.array([used, [0, 0, 0]], dtype=np.int64)
    ids = [node_name, "machine2" if node_name != "machine2" else "other"]
    ask_node = -1 if pod_node is None else (ids.index(pod_node) if pod_node in ids else 1)
    z = np.zeros(1, dtype=np.uint64)
    return synth._finish("pair", D, 0, total, avail, np.zeros(2, np.uint64), np.zeros(2, np.uint64), ids,
                         synth._single_queue(D), np.ones(1, np.int32), np.zeros(1, np.int32),
                         np.array([req], dtype=np.int64), z, z.copy(), z.copy(), ask_node=np.array([ask_node], np.int32))


def test_reference_general_predicates_table(oracle):
    cases = json.load(open(os.path.join(GOLD, "general_predicates.json")))["cases"]
    seen = 0
    for c in cases:
        if c.get("slow_path"):
            continue   # NodePorts is not a bitmask predicate: the ask would be flagged slow-path
        s = one_pair_snapshot(c["alloc"], c["used"], c["req"], c["pod_node"], c["node"])
        rc = oracle.predicate(s, 0, 0)
        if c["req"] == [0, 0, 1] or any(c["req"]):
            assert (rc == 0) == c["fits"], (c["name"], rc)
        if not c["fits"]:
            want = {"NodeResourcesFit": (4, 8), "NodeName": (5,)}[c["plugin"]]
            assert rc in want, (c["name"], rc)
        seen += 1
    assert seen >= 6


def test_node_score_hand_computed(oracle):
    w = [1.0, 1.0, 0.0, 0.0]
    tot = [32000, 256 << 30, 110, 0]
    assert oracle.node_score(0, w, tot, tot) == 0.0
    av = [16000, 128 << 30, 100, 0]                       # half used on both weighted dims
    assert oracle.node_score(0, w, tot, av) == 0.5
    assert oracle.node_score(1, w, tot, av) == 0.5        # binpacking = 1 - fair
    av = [31900, (256 << 30) - (128 << 20), 109, 0]       # one 100m/128Mi pod
    want = ((1.0 - 31900 / 32000) + (1.0 - ((256 << 30) - (128 << 20)) / (256 << 30))) / 2.0
    assert oracle.node_score(0, w, tot, av) == want
    assert oracle.node_score(0, [0, 0, 0, 0], tot, av) == 0.0      # no weighted type: 0
    assert oracle.node_score(0, [1.0, 1.0, 0, 0], [0, 0, 110, 0], [0, 0, 5, 0]) == 0.0   # 0/0 NaN shares skipped
    assert oracle.node_score(0, w, tot, [-1000, 256 << 30, 110, 0]) == (1.0 - (-1000 / 32000)) / 2.0   # over-committed


def test_drf_known_answer_from_resource_fairness_e2e(oracle):
    """/root/reference/test/e2e/resource_fairness/resource_fairness_test.go:104-109,170: apps holding (cpu,mem)
    shares (0.3,0.1) (0.2,0.05) (0.1,0.15) are served app2, app1, app0 = ascending dominant share."""
    g = [1000, 1000]
    a0, a1, a2 = [300, 100], [200, 50], [100, 150]
    assert oracle.comp_usage_ratio_separately(a2, g, a1, g) == -1
    assert oracle.comp_usage_ratio_separately(a1, g, a0, g) == -1
    assert oracle.comp_usage_ratio_separately(a0, g, a2, g) == 1
    assert oracle.comp_usage_ratio_separately(a0, g, a0, g) == 0
    assert oracle.comp_usage_ratio_separately([0, 0], g, [0, 1], g) == -1      # nothing allocated sorts first
    assert oracle.comp_usage_ratio_separately([5, 0], [-1, -1], [4, 0], [-1, -1]) == 1   # no guarantee: raw usage


def test_node_id_tie_break_is_bytewise_string_order(oracle):
    s = synth.kwok(12, 1, 12, variant="sized")    # equal empty nodes: all scores 0 -> NodeID decides; a commit raises the score
    r = oracle.run(s)
    picked = [s.node_id[n] for n in r["node"]]
    assert picked == sorted(s.node_id, key=lambda x: x.encode())
    assert picked[:3] == ["kwok-node-0", "kwok-node-1", "kwok-node-10"]   # "…-10" < "…-2"
    bare = synth.kwok(12, 1, 12, variant="bare")  # {pods:1} asks do not move the (vcore,memory) score: all stack on the first NodeID
    assert set(oracle.run(bare)["node"]) == {0}


@pytest.mark.parametrize("make", [
    lambda: synth.kwok(30, 5, 20), lambda: synth.kwok(30, 5, 20, variant="bare", policy=1),
    lambda: synth.perf(80, 8, 25), lambda: synth.perf(80, 8, 25, masks=True),
    lambda: synth.perf(80, 8, 25, masks=True, policy=1), lambda: synth.perf(6, 4, 200),
    lambda: synth.hier(60, 3, 3, 2, 12, priorities=True), lambda: synth.hier(60, 2, 4, 2, 12, masks=True, quota_frac=2.0),
    lambda: synth.hier(40, 2, 2, 3, 12, priorities=True, leaf_sort=1), lambda: synth.gangs(30, 12, 5), lambda: synth.gangs(40, 20, 4, fill=1.6, policy=1),
])
def test_two_independent_restatements_agree(oracle, make):
    s = make()
    r, p = oracle.run(s), py_oracle.run(s)
    assert list(r["ask"]) == p["ask"]
    assert list(r["node"]) == p["node"]
    assert list(r["state"]) == p["state"]
    assert r["avail"].tolist() == p["avail"]


def test_edge_cases(oracle):
    s = synth.perf(5, 2, 10)
    s.node_flags[:] = 0                                   # no schedulable node: everything NOFIT
    r = oracle.run(s)
    assert len(r["ask"]) == 0 and set(r["state"]) == {2}
    s = synth.perf(5, 2, 10)
    s.ask_req[:] = 0                                      # nothing requested: invalid (preAllocateCheck)
    assert set(oracle.run(s)["state"]) == {5}
    s = synth.perf(5, 2, 10)
    s.ask_flags[::2] = 1                                  # slow-path asks are never bound by the fast path
    r = oracle.run(s)
    assert set(r["state"][::2]) == {4} and set(r["state"][1::2]) == {1}
    s = synth.perf(5, 2, 10)
    s.ask_node[:] = 3                                     # pod.Spec.NodeName
    assert set(oracle.run(s)["node"]) == {3}
    empty = synth.perf(5, 1, 0)
    assert len(oracle.run(empty)["ask"]) == 0


def test_frozen_oracle_goldens(oracle):
    gold = json.load(open(os.path.join(GOLD, "oracle_bindings.json")))
    for c in gold["cases"]:
        s = getattr(synth, c["generator"])(**c["args"])
        r = oracle.run(s)
        assert len(r["ask"]) == c["n_bindings"]
        assert f"{oracle.bindings_hash(r['ask'], r['node']):#x}" == c["hash"], c["args"]
        if "ask" in c:
            assert [int(a) for a in r["ask"]] == c["ask"]
            assert [s.node_id[n] for n in r["node"]] == c["node_id_of_binding"]


def test_priority_order_across_sibling_queues_from_priority_scheduling_e2e(oracle):
    """/root/reference/test/e2e/priority_scheduling/priority_scheduling_test.go:70-133,407-476: low, normal and high
    priority pods sit in two sibling queues under a parent whose quota admits one pod at a time; whenever room appears
    the reference starts high, then normal, then low -- regardless of submit order and of which queue they are in."""
    from oracle import py_oracle
    from yunikorn_k8shim_b200 import synth
    done = []
    for expect in ("high", "normal", "low"):
        s = synth.priority_fence(quota_pods=1, done=done)
        for run in (oracle.run, py_oracle.run):
            r = run(s)
            assert [s.meta["apps"][a] for a in r["ask"]] == [expect]            # one pod fits the quota: it must be this one
            assert sorted(int(x) for x in r["state"]) == [1] + [3] * (len(s.meta["apps"]) - 1)   # the others: headroom
        done.append(expect)
    # with room for all three the cycle serves them in priority order
    s = synth.priority_fence(quota_pods=3)
    for run in (oracle.run, py_oracle.run):
        assert [s.meta["apps"][a] for a in run(s)["ask"]] == ["high", "normal", "low"]
    # priority_scheduling_test.go:179-251: the same order when it comes from queue offsets (+100 / 0 / -100) instead of
    # pod priority classes
    done = []
    for expect in ("high", "normal", "low"):
        s = synth.priority_offsets(quota_pods=1, done=done)
        for run in (oracle.run, py_oracle.run):
            assert [s.meta["apps"][a] for a in run(s)["ask"]] == [expect]
        done.append(expect)
    s = synth.priority_offsets(quota_pods=3)
    for run in (oracle.run, py_oracle.run):
        assert [s.meta["apps"][a] for a in run(s)["ask"]] == ["high", "normal", "low"]


def _fence_snapshot(fenced):
    from yunikorn_k8shim_b200 import synth
    D = 4
    base = synth.perf(4, 1, 1)
    qp = np.array([-1, 0, 0], dtype=np.int32)                                # root -> A, B
    unset = np.full((3, D), -1, dtype=np.int64)
    queues = (qp, unset.copy(), unset.copy(), np.zeros((3, D), dtype=np.int64), np.zeros(3, dtype=np.uint8))
    req = np.tile(np.array([100, 100 * 1000 * 1000, 1, 0], dtype=np.int64), (4, 1))
    z = np.zeros(4, dtype=np.uint64)
    s = synth._finish("fence", D, synth.POLICY_FAIR, base.node_total, base.node_avail, base.node_taint, base.node_label, base.node_id,
                      queues, np.array([1, 2], dtype=np.int32), np.array([0, 0, 1, 1], dtype=np.int32), req, z, z.copy(), z.copy(),
                      ask_prio=np.array([-100, -100, 0, 0], dtype=np.int32))
    s.q_prio_fence = np.array([0, 1 if fenced else 0, 0], dtype=np.uint8)
    s.q_prio_offset = np.zeros(3, dtype=np.int32)
    return s


def test_fence_hides_the_priorities_below_it(oracle):
    """priority.policy = fence [EXT yunikorn-core, restated; unpinned by a reference test]: queue A holds two pods of
    priority -100, queue B two of priority 0.  Unfenced, B's pods go first (queue priority 0 > -100).  With A fenced its
    parent sees only A's offset (0): the queues tie on priority and alternate by share, starting with A (index order)."""
    from oracle import py_oracle
    for fenced, want in ((False, [2, 3, 0, 1]), (True, [0, 2, 1, 3])):
        s = _fence_snapshot(fenced)
        for run in (oracle.run, py_oracle.run):
            assert list(run(s)["ask"]) == want, fenced
