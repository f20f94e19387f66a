"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded snapshots.
Bit-exact bar: identical (ask,node) bindings in identical commit order, identical final availability."""
import numpy as np
import pytest

from yunikorn_k8shim_b200 import Engine, synth

pytestmark = pytest.mark.gpu


def _check(snap, oracle, **kw):
    want = oracle.run(snap)
    with Engine.for_snapshot(snap, **kw) as e:
        ask, node, _ = e.cycle(snap.n_asks)
        avail = e.nodes_available(np.arange(snap.n_nodes))
        states = e.ask_states(np.arange(snap.n_asks))
        st = e.stats()
    assert len(ask) == len(want["ask"])
    assert np.array_equal(ask, want["ask"]), "ask order differs"
    assert np.array_equal(node, want["node"]), "node choice differs"
    assert np.array_equal(avail, want["avail"])
    assert np.array_equal(states, want["state"])
    assert st["sweep_launches"] > 0 and st["evaluations"] > 0
    return st


@pytest.mark.parametrize("variant", ["bare", "sized"])
@pytest.mark.parametrize("policy", [synth.POLICY_FAIR, synth.POLICY_BINPACKING])
def test_config1_kwok(oracle, variant, policy):
    _check(synth.kwok(100, 10, 50, variant=variant, policy=policy), oracle, batch=128)


@pytest.mark.parametrize("batch", [64, 1000, 4096])
def test_config2_small(oracle, batch):
    _check(synth.perf(700, 20, 100), oracle, batch=batch)


@pytest.mark.parametrize("policy", [synth.POLICY_FAIR, synth.POLICY_BINPACKING])
def test_config3_small(oracle, policy):
    _check(synth.perf(900, 20, 100, masks=True, policy=policy), oracle, batch=512)


def test_config2_full(oracle):
    _check(synth.perf(), oracle, batch=2048)


def test_config3_full(oracle):
    _check(synth.perf(masks=True), oracle, batch=2048)


def test_overcommitted_cluster(oracle):
    # far more demand than capacity: many asks must end NOFIT, in the oracle's order
    snap = synth.perf(8, 10, 200, masks=True)
    st = _check(snap, oracle, batch=256)
    assert st["nofit"] > 0
