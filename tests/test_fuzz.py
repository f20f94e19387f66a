"""Randomized parity over snapshots that mix every feature of the path (synth.fuzz):
CPU: the two independent oracle restatements must agree; GPU: the engine must match the oracle bit for bit."""
import numpy as np
import pytest

from yunikorn_k8shim_b200 import synth
from oracle import py_oracle


@pytest.mark.parametrize("seed", range(60))
def test_oracles_agree_on_fuzz(oracle, seed):
    s = synth.fuzz(seed)
    r, p = oracle.run(s), py_oracle.run(s)
    assert list(r["ask"]) == p["ask"], s.name
    assert list(r["node"]) == p["node"], s.name
    assert list(r["state"]) == p["state"], s.name
    if seed % 5 == 0:
        k = max(1, len(p["ask"]) // 2)
        r2, p2 = oracle.run(s, max_bindings=k), py_oracle.run(s, max_bindings=k)
        assert list(r2["ask"]) == p2["ask"] and list(r2["node"]) == p2["node"]


def test_fuzz_exercises_the_features(oracle):
    seen = np.zeros(6, dtype=int)
    gangs = fair = binp = 0
    for seed in range(60):
        s = synth.fuzz(seed)
        seen += np.bincount(oracle.run(s)["state"], minlength=6)
        gangs += int((s.ask_gang >= 0).any())
        fair += int(s.q_sort.any())
        binp += s.policy
    assert (seen[1:] > 0).all(), seen      # allocated, nofit, skipped, slowpath, invalid all occur
    assert gangs > 10 and fair > 10 and 10 < binp < 50


@pytest.mark.gpu
@pytest.mark.parametrize("batch,share,commit", [(7, True, "host"), (64, True, "host"), (1024, True, "host"), (64, False, "host"),
                                                (7, True, "device"), (64, True, "device"), (1024, True, "device")])
def test_engine_matches_oracle_on_fuzz(oracle, batch, share, commit):
    from yunikorn_k8shim_b200 import Engine
    for seed in range(60 if commit == "host" else 120):
        s = synth.fuzz(seed)
        want = oracle.run(s)
        with Engine.for_snapshot(s, batch=batch, share_rows=share, commit=commit) as e:
            try:
                ask, node, _ = e.cycle(s.n_asks)
            except Exception as exc:          # a gang larger than a tiny batch is a documented error, not a mismatch
                assert batch == 7 and "gang" in str(exc), (seed, exc)
                continue
            states = e.ask_states(np.arange(s.n_asks))
            avail = e.nodes_available(np.arange(s.n_nodes))
        assert np.array_equal(ask, want["ask"]), (seed, batch)
        assert np.array_equal(node, want["node"]), (seed, batch)
        assert np.array_equal(states, want["state"]), (seed, batch)
        assert np.array_equal(avail, want["avail"]), (seed, batch)


@pytest.mark.gpu
@pytest.mark.parametrize("batch,uniform_min", [(64, 2), (1024, 3), (7, 2)])
def test_engine_uniform_runs_match_oracle_on_fuzz(oracle, monkeypatch, batch, uniform_min):
    """asks that come in runs of identical (request, signature): the device commit decides every run of at least
    uniform_min entries by the grid-wide sort of csrc/yk_uniform.cuh instead of the windowed kernel"""
    from yunikorn_k8shim_b200 import Engine
    monkeypatch.setenv("YK_UNIFORM_MIN", str(uniform_min))
    decided = 0
    for seed in range(120):
        s = synth.runny(synth.fuzz(seed), seed)
        want = oracle.run(s)
        with Engine.for_snapshot(s, batch=batch, commit="device") as e:
            try:
                ask, node, _ = e.cycle(s.n_asks)
            except Exception as exc:
                assert batch == 7 and "gang" in str(exc), (seed, exc)
                continue
            states = e.ask_states(np.arange(s.n_asks))
            avail = e.nodes_available(np.arange(s.n_nodes))
            decided += e.stats()["uniform_asks"]
        assert np.array_equal(ask, want["ask"]), (seed, batch)
        assert np.array_equal(node, want["node"]), (seed, batch)
        assert np.array_equal(states, want["state"]), (seed, batch)
        assert np.array_equal(avail, want["avail"]), (seed, batch)
    assert decided > 200


@pytest.mark.gpu
@pytest.mark.parametrize("policy", [synth.POLICY_FAIR, synth.POLICY_BINPACKING])
def test_engine_gang_rollback_stress(oracle, policy):
    """placed-then-undone gangs interleaved with plain asks (synth.poisoned_gangs), one long epoch"""
    from yunikorn_k8shim_b200 import Engine
    for seed in range(20):
        s = synth.poisoned_gangs(seed, policy=policy)
        want = oracle.run(s)
        for batch, commit in ((16, "host"), (1024, "host"), (64, "device")):
            with Engine.for_snapshot(s, batch=batch, commit=commit) as e:
                ask, node, _ = e.cycle(s.n_asks)
                avail = e.nodes_available(np.arange(s.n_nodes))
            assert np.array_equal(ask, want["ask"]), (seed, batch)
            assert np.array_equal(node, want["node"]), (seed, batch)
            assert np.array_equal(avail, want["avail"]), (seed, batch)
