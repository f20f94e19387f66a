"""Preemption victim search (PreemptionPredicates, predicate_manager.go:137-175): the reference's known answers
on the oracle (CPU) and on the device kernel (GPU), plus seeded parity of the batched device search."""
import json
import os

import numpy as np
import pytest

from yunikorn_k8shim_b200 import synth

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preemption_predicates.json")))


def golden_snapshot(req, with_victims):
    D = 3
    alloc = np.array([GOLD["node_alloc"]], dtype=np.int64)
    used = np.sum(np.array(GOLD["victims"], dtype=np.int64), axis=0) if with_victims else np.zeros(3, dtype=np.int64)
    z = np.zeros(1, dtype=np.uint64)
    return synth._finish("preempt", D, 0, alloc, alloc - used, z, z.copy(), ["node0"], synth._single_queue(D),
                         np.ones(1, np.int32), np.zeros(1, np.int32), np.array([req], dtype=np.int64), z.copy(), z.copy(), z.copy())


@pytest.mark.parametrize("c", GOLD["cases"], ids=[c["name"][:30] for c in GOLD["cases"]])
def test_reference_known_answers_on_oracle(oracle, c):
    s = golden_snapshot(c["req"], c["victims"])
    v = np.array(GOLD["victims"] if c["victims"] else np.zeros((0, 3)), dtype=np.int64)
    assert oracle.preemption_index(s, 0, 0, v, c["start"]) == c["index"]


@pytest.mark.gpu
def test_reference_known_answers_on_device():
    from yunikorn_k8shim_b200 import Engine
    for c in GOLD["cases"]:
        s = golden_snapshot(c["req"], c["victims"])
        v = np.array(GOLD["victims"] if c["victims"] else np.zeros((0, 3)), dtype=np.int64)
        with Engine.for_snapshot(s) as e:
            assert e.preemption_search([0], [0], [v], [c["start"]])[0] == c["index"], c["name"]


@pytest.mark.gpu
def test_batched_search_matches_oracle(oracle):
    from yunikorn_k8shim_b200 import Engine
    s = synth.perf(64, 4, 40, masks=True, seed=31)
    s.node_avail[:, 0] = s.node_avail[:, 0] // 50              # nearly full nodes: victims are needed
    s.node_avail[:, 1] = s.node_avail[:, 1] // 50
    rng = np.random.default_rng(5)
    asks = rng.integers(0, s.n_asks, 300)
    nodes = rng.integers(0, s.n_nodes, 300)
    victims, starts = [], []
    for q in range(300):
        nv = int(rng.integers(0, 70))                          # crosses the 32-victim warp step
        v = np.zeros((nv, s.D), dtype=np.int64)
        v[:, 0] = rng.integers(0, 800, nv)
        v[:, 1] = rng.integers(0, 4 << 30, nv)
        v[:, 2] = 1
        victims.append(v)
        starts.append(int(rng.integers(0, nv + 2)))
    want = [oracle.preemption_index(s, int(a), int(n), v, st) for a, n, v, st in zip(asks, nodes, victims, starts)]
    with Engine.for_snapshot(s) as e:
        got = e.preemption_search(asks, nodes, victims, starts)
    assert got.tolist() == want
    assert any(w >= 0 for w in want) and any(w < 0 for w in want)
