"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded snapshots.
Bit-exact bar: identical (ask,node) bindings in identical commit order, identical final availability."""
import numpy as np
import pytest

from yunikorn_k8shim_b200 import Engine, synth

pytestmark = pytest.mark.gpu


def _run(snap, want, **kw):
    with Engine.for_snapshot(snap, **kw) as e:
        ask, node, _ = e.cycle(snap.n_asks)
        avail = e.nodes_available(np.arange(snap.n_nodes))
        states = e.ask_states(np.arange(snap.n_asks))
        st = e.stats()
    assert len(ask) == len(want["ask"]), kw
    assert np.array_equal(ask, want["ask"]), ("ask order differs", kw)
    assert np.array_equal(node, want["node"]), ("node choice differs", kw)
    assert np.array_equal(avail, want["avail"]), kw
    assert np.array_equal(states, want["state"]), kw
    return st


def _check(snap, oracle, **kw):
    """both commits: the device-resident one (wherever the cycle is eligible for it) and sweep + host commit"""
    want = oracle.run(snap)
    dev = _run(snap, want, commit="device", **kw)
    assert dev["sweep_launches"] + dev["lattice_launches"] > 0 or snap.n_nodes == 0
    st = _run(snap, want, commit="host", **kw)
    assert st["sweep_launches"] > 0 and st["evaluations"] > 0 and st["lattice_launches"] == 0
    st["device"] = dev
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
    st = _check(synth.perf(), oracle, batch=2048)
    assert st["rows_swept"] * 50 < st["asks_swept"]      # a few dozen distinct pod shapes: rows are shared
    dev = st["device"]                                    # the device commit took the whole cycle: no sweep, no hand-over
    assert dev["lattice_asks"] == 50_000 and dev["sweep_launches"] == 0 and dev["lattice_handoffs"] == 0
    assert dev["lattice_subruns"] < 1000




def test_config2_full_one_row_per_ask(oracle):
    st = _check(synth.perf(), oracle, batch=2048, share_rows=False)
    assert st["rows_swept"] == st["asks_swept"] >= 50_000


def test_config3_full(oracle):
    st = _check(synth.perf(masks=True), oracle, batch=2048)
    assert st["device"]["lattice_asks"] == 50_000 and st["device"]["sweep_launches"] == 0


def test_reference_benchmark_shape_full(oracle):
    """the reference's own benchmark shape: every pod identical (pkg/shim/scheduler_perf_test.go:283-288)"""
    st = _check(synth.reference_shape(), oracle, batch=4096)
    assert st["device"]["lattice_asks"] == 50_000
    assert st["device"]["uniform_asks"] == 50_000 and st["device"]["lattice_subruns"] == 0   # one uniform run: no chain at all


def test_default_commit_choice(oracle):
    """the engine's default (auto): cycles made of long uniform runs are decided on the device, everything else by the sweep +
    host commit -- same bindings either way"""
    snap = synth.reference_shape(2_000, 40, 125)                     # one run of 5 000 identical asks
    st = _run(snap, oracle.run(snap), batch=4096)
    assert st["lattice_cycles"] == 1 and st["uniform_asks"] == 5_000 and st["sweep_launches"] == 0
    snap = synth.perf(2_000, 40, 125)                                # every ask draws its own request class: no runs
    st = _run(snap, oracle.run(snap), batch=4096)
    assert st["lattice_cycles"] == 0 and st["sweep_launches"] > 0
    snap = synth.reference_shape(2_000, 40, 125, policy=synth.POLICY_BINPACKING)   # not eligible: binpacking node sort
    st = _run(snap, oracle.run(snap), batch=4096)
    assert st["lattice_cycles"] == 0 and st["sweep_launches"] > 0
    snap = synth.kwok(2_000, 20, 250, variant="bare")                # kwok-perf-test pods: {pods: 1} only, keys never move
    st = _run(snap, oracle.run(snap), batch=4096)
    assert st["lattice_cycles"] == 1 and st["uniform_asks"] == snap.n_asks and st["uniform_retries"] == 0
    snap = synth.reference_shape(300, 8, 125)                        # 1 000 asks: below the shortest run worth a device pass
    st = _run(snap, oracle.run(snap), batch=4096)
    assert st["lattice_cycles"] == 0


def test_uniform_runs_inside_mixed_batches(oracle, monkeypatch):
    """config-2/3/4-sized snapshots whose asks come in runs: uniform runs and windowed stretches alternate inside one batch,
    on the same device-resident node state"""
    monkeypatch.setenv("YK_UNIFORM_MIN", "6")
    uniform = 0
    for snap, batch in ((synth.runny(synth.perf(1500, 20, 100), 1), 4096), (synth.runny(synth.perf(1500, 20, 100, masks=True), 2), 4096),
                        (synth.runny(synth.hier(400, 3, 3, 2, 40), 3), 512), (synth.runny(synth.gangs(300, 60, 5), 4), 1000)):
        want = oracle.run(snap)
        st = _run(snap, want, commit="device", batch=batch)
        assert st["lattice_subruns"] > 0, snap.name
        uniform += st["uniform_asks"]
    assert uniform > 1000
    # more identical pods than the cluster holds
    snap = synth.reference_shape(6, 3, 300)
    snap.node_total[:, 2] = snap.node_avail[:, 2] = 110
    st = _run(snap, oracle.run(snap), commit="device", batch=4096)
    assert st["allocations"] == 660 and st["uniform_asks"] > 0


def test_overcommitted_cluster(oracle):
    # far more demand than capacity: many asks must end NOFIT, in the oracle's order
    snap = synth.perf(8, 10, 200, masks=True)
    st = _check(snap, oracle, batch=256)
    assert st["nofit"] > 0


# ---- config 4: hierarchical queues, DRF parents, quotas, priorities (placement-sensitive order) ----
@pytest.mark.parametrize("prio", [False, True])
def test_config4_small(oracle, prio):
    st = _check(synth.hier(300, 3, 4, 2, 40, masks=True, priorities=prio, seed=21), oracle, batch=128)
    assert st["skipped"] > 0


def test_config4_overcommitted_rewinds(oracle):
    # too few nodes: placement failures inside a DRF order force rewinds
    st = _check(synth.hier(6, 3, 3, 2, 60, priorities=True, seed=22), oracle, batch=64)
    assert st["nofit"] > 0


def test_config4_full(oracle):
    _check(synth.hier(), oracle, batch=4096)


def test_fair_leaf_policy(oracle):
    _check(synth.hier(200, 2, 3, 3, 30, masks=True, priorities=True, seed=23, leaf_sort=synth.SORT_FAIR), oracle, batch=64)


# ---- config 5: gangs, all-or-nothing ----
@pytest.mark.parametrize("policy", [synth.POLICY_FAIR, synth.POLICY_BINPACKING])
def test_config5_small(oracle, policy):
    st = _check(synth.gangs(60, 40, 5, fill=1.4, policy=policy), oracle, batch=64)
    assert st["nofit"] > 0


def test_config5_gangs_in_drf_tree(oracle):
    s = synth.hier(30, 2, 3, 2, 24, seed=9)
    s.ask_gang[:] = np.arange(s.n_asks) // 4
    _check(s, oracle, batch=32)


def test_config5_full(oracle):
    st = _check(synth.gangs(), oracle, batch=4096)
    assert st["nofit"] > 0
    # the device commit places the gangs until the cluster is full; the first gang that has to be rolled back hands the
    # rest of the cycle to the host commit
    assert st["device"]["lattice_asks"] > 5_000 and st["device"]["lattice_handoffs"] == 1


def test_gang_larger_than_batch_is_an_error(oracle):
    from yunikorn_k8shim_b200 import YkError
    s = synth.gangs(30, 4, 20)
    with Engine.for_snapshot(s, batch=8) as e:
        with pytest.raises(YkError):
            e.cycle(s.n_asks)


# ---- edge cases the reference tests at this boundary ----
def test_edges(oracle):
    s = synth.perf(5, 2, 10)
    s.node_flags[:] = 0
    _check(s, oracle, batch=16)                                   # no schedulable node
    s = synth.perf(5, 2, 10)
    s.ask_req[::3] = 0
    _check(s, oracle, batch=16)                                   # invalid requests
    s = synth.perf(5, 2, 10)
    s.ask_flags[::2] = 1
    _check(s, oracle, batch=16)                                   # slow-path asks are left alone
    s = synth.perf(9, 2, 10)
    s.ask_node[:] = 3
    _check(s, oracle, batch=16)                                   # pod.Spec.NodeName
    s = synth.perf(9, 2, 10)
    s.node_flags[4] = synth.NODE_SCHEDULABLE | synth.NODE_RESERVED
    s.node_avail[2, 0] = -5000                                    # over-committed node: FitIn clamps at 0
    _check(s, oracle, batch=16)


def test_max_bindings_and_second_cycle(oracle):
    s = synth.perf(50, 4, 50, masks=True)
    want = oracle.run(s, max_bindings=77)
    full = oracle.run(s)
    for commit in ("device", "host"):
        with Engine.for_snapshot(s, batch=32, commit=commit) as e:
            ask, node, _ = e.cycle(77)
            assert np.array_equal(ask, want["ask"]) and np.array_equal(node, want["node"])
            ask2, node2, _ = e.cycle(s.n_asks)                        # the rest, on the state the first cycle left
            assert np.array_equal(np.concatenate([ask, ask2]), full["ask"])
            assert np.array_equal(np.concatenate([node, node2]), full["node"])


def test_two_uniform_cycles_on_the_device(oracle):
    """the state a device-committed cycle leaves (host and device node tables, ask states, queue accounting) is what the next
    cycle starts from: two cycles, both decided as uniform runs on the device, equal one oracle pass"""
    s = synth.reference_shape(500, 96, 125)                              # 12 000 identical asks
    full = oracle.run(s)
    with Engine.for_snapshot(s, batch=4096) as e:
        ask, node, _ = e.cycle(6000)
        ask2, node2, _ = e.cycle(s.n_asks)
        st = e.stats()
        assert np.array_equal(np.concatenate([ask, ask2]), full["ask"])
        assert np.array_equal(np.concatenate([node, node2]), full["node"])
        assert np.array_equal(e.nodes_available(np.arange(s.n_nodes)), full["avail"])
        assert st["lattice_cycles"] == 2 and st["uniform_asks"] == s.n_asks
        e.release(np.concatenate([ask, ask2]))
        assert np.array_equal(e.nodes_available(np.arange(s.n_nodes)), s.node_avail)


def test_release_gives_resources_back(oracle):
    s = synth.perf(6, 2, 300)                                     # overcommitted
    want = oracle.run(s)
    for commit in ("device", "host"):
        with Engine.for_snapshot(s, batch=64, commit=commit) as e:
            ask, node, _ = e.cycle(s.n_asks)
            assert np.array_equal(ask, want["ask"])
            e.release(ask)
            assert np.array_equal(e.nodes_available(np.arange(s.n_nodes)), s.node_avail)


def test_device_score_and_evaluate_match_oracle(oracle):
    s = synth.perf(200, 4, 20, masks=True)
    s.node_avail[:, 0] -= np.arange(200) * 37
    with Engine.for_snapshot(s) as e:
        sc = e.node_scores(np.arange(s.n_nodes))
        want = [oracle.node_score(s.policy, s.weights, s.node_total[n], s.node_avail[n]) for n in range(s.n_nodes)]
        assert sc.tolist() == want                                # float64 bit-exact
        for a in range(0, s.n_asks, 7):
            for n in range(0, s.n_nodes, 11):
                got, exp = e.evaluate(a, n), oracle.predicate(s, a, n)
                assert (got == 0) == (exp == 0) and (got == exp or {got, exp} <= {4, 8}), (a, n, got, exp)


def _sub_snapshot(s, keep_nodes, keep_asks):
    """oracle-side view of a snapshot after some nodes / asks were removed (indices are preserved by the engine,
    so the oracle gets the same arrays with the removed nodes unschedulable+full and the removed asks slow-path)"""
    import copy
    t = copy.deepcopy(s)
    gone_n = np.setdiff1d(np.arange(s.n_nodes), keep_nodes)
    t.node_flags[gone_n] = 0
    gone_a = np.setdiff1d(np.arange(s.n_asks), keep_asks)
    t.ask_flags[gone_a] = 1
    return t


def test_remove_nodes_and_asks_then_cycle(oracle):
    s = synth.perf(120, 6, 40, masks=True, seed=41)
    rng = np.random.default_rng(2)
    keep_n = np.sort(rng.choice(s.n_nodes, 90, replace=False))
    keep_a = np.sort(rng.choice(s.n_asks, 200, replace=False))
    want = oracle.run(_sub_snapshot(s, keep_n, keep_a))
    with Engine.for_snapshot(s, batch=64) as e:
        e.nodes_remove(np.setdiff1d(np.arange(s.n_nodes), keep_n))
        e.asks_remove(np.setdiff1d(np.arange(s.n_asks), keep_a))
        ask, node, _ = e.cycle(s.n_asks)
        assert np.array_equal(ask, want["ask"]) and np.array_equal(node, want["node"])
        assert set(node.tolist()) <= set(keep_n.tolist())
        # removed asks report ABSENT, removed nodes are gone from evaluation
        assert set(e.ask_states(np.setdiff1d(np.arange(s.n_asks), keep_a)).tolist()) == {255}
        assert e.evaluate(int(keep_a[0]), int(np.setdiff1d(np.arange(s.n_nodes), keep_n)[0])) == 9


def test_node_update_between_cycles(oracle):
    """capacity / availability updates (SchedulerAPI.UpdateNode UPDATE, cordon) arriving between two cycles"""
    import copy
    s = synth.perf(60, 4, 60, seed=43)
    first = oracle.run(s, max_bindings=100)
    with Engine.for_snapshot(s, batch=32) as e:
        ask, node, _ = e.cycle(100)
        assert np.array_equal(ask, first["ask"]) and np.array_equal(node, first["node"])
        # the world after cycle 1, then: node 5 cordoned, node 7 grows
        t = copy.deepcopy(s)
        t.node_avail = e.nodes_available(np.arange(s.n_nodes))
        t.node_flags[5] = 0
        t.node_total[7, 0] += 16000
        t.node_avail[7, 0] += 16000
        e.nodes_upsert([5, 7], t.node_total[[5, 7]], t.node_avail[[5, 7]], t.node_taint[[5, 7]], t.node_label[[5, 7]],
                       s.node_rank()[[5, 7]], t.node_flags[[5, 7]])
        t.ask_flags[first["ask"]] = 1                      # already bound: the oracle must leave them alone
        want = oracle.run(t)
        ask2, node2, _ = e.cycle(s.n_asks)
        assert np.array_equal(ask2, want["ask"]) and np.array_equal(node2, want["node"])
        assert 5 not in set(node2.tolist())


def test_error_paths_and_empty_cluster():
    from yunikorn_k8shim_b200 import YkError
    s = synth.perf(10, 2, 10)
    # no nodes at all: every ask ends NOFIT, no crash
    e = Engine(D=s.D, max_nodes=16, max_asks=64, max_apps=4, max_queues=4)
    e.queues_set(s.q_parent, s.q_guaranteed, s.q_max, s.q_alloc, s.q_sort)
    e.apps_upsert(np.arange(s.n_apps), s.app_queue, s.app_submit)
    e.asks_upsert(np.arange(s.n_asks), s.ask_req, s.ask_app, s.ask_create)
    ask, node, _ = e.cycle(100)
    assert len(ask) == 0 and set(e.ask_states(np.arange(s.n_asks)).tolist()) == {2}
    # argument errors come back as codes, never as a crash
    with pytest.raises(YkError) as ei:
        e.nodes_upsert([99], s.node_total[:1], s.node_avail[:1], name_rank=[0])
    assert ei.value.code == -1
    with pytest.raises(YkError):
        e.asks_upsert([0], s.ask_req[:1], [3], [0])            # unknown application
    with pytest.raises(YkError):
        e.release([0])                                          # holds no allocation
    # a node whose weighted totals are zero scores 0 (types absent from total are skipped): still schedulable order
    tot, av = s.node_total[:2].copy(), s.node_avail[:2].copy()
    tot[0, 0], av[0, 0] = 0, 0
    tot[0, 1], av[0, 1] = 0, 0
    e.nodes_upsert([0, 1], tot, av, name_rank=[0, 1])
    assert e.node_scores([0])[0] == 0.0
    e.close()


def test_queue_priority_properties(oracle):
    """yk_queues_priority (priority.offset / priority.policy = fence) through the C ABI: the offsets scenario of
    priority_scheduling_test.go:179-251, the pod-priority scenario of :70-133, and a fenced queue"""
    done = []
    for expect in ("high", "normal", "low"):
        s = synth.priority_offsets(quota_pods=1, done=done)
        _check(s, oracle, batch=8)
        with Engine.for_snapshot(s, batch=8) as e:
            ask, node, _ = e.cycle(s.n_asks)
        assert [s.meta["apps"][a] for a in ask] == [expect]
        done.append(expect)
    _check(synth.priority_offsets(quota_pods=3), oracle, batch=2)
    _check(synth.priority_fence(quota_pods=3), oracle, batch=2)
    s = synth.hier(40, 2, 3, 2, 12, priorities=True, seed=31)
    s.q_prio_offset = (np.arange(s.n_queues) % 3 * 50 - 50).astype(np.int32)
    s.q_prio_fence = (np.arange(s.n_queues) % 4 == 1).astype(np.uint8)
    _check(s, oracle, batch=16)


def test_reupserted_asks_get_fresh_signature_numbers(oracle):
    """signature numbers are handed out when asks are upserted and persist across cycles: overwriting asks -- also ones
    that represent a number -- with other predicate inputs must never merge them with the old signature's row"""
    import copy
    s = synth.perf(80, 4, 60, masks=True, seed=51)
    s.ask_tol[:] = s.ask_tol[0]
    s.ask_need[:] = 0
    s.ask_deny[:] = 0                                    # few signatures: the epoch-rows path
    first = oracle.run(s, max_bindings=50)
    for commit in ("host", "device"):
        with Engine.for_snapshot(s, batch=32, commit=commit) as e:
            ask, node, _ = e.cycle(50)
            assert np.array_equal(ask, first["ask"]) and np.array_equal(node, first["node"])
            t = copy.deepcopy(s)
            t.node_avail = e.nodes_available(np.arange(s.n_nodes))
            t.ask_flags[first["ask"]] = 1                # bound: the oracle leaves them alone
            pend = np.setdiff1d(np.arange(s.n_asks), first["ask"])
            chg = pend[::3]                              # a third of the pending asks change their predicate inputs
            t.ask_tol[chg] = np.uint64(0xFFFF)
            t.ask_need[chg] = np.uint64(1) << (np.arange(len(chg)) % 8).astype(np.uint64)
            t.ask_req[chg[::2], 0] += 30
            e.asks_upsert(chg, t.ask_req[chg], t.ask_app[chg], t.ask_create[chg], t.ask_tol[chg], t.ask_need[chg], t.ask_deny[chg])
            want = oracle.run(t)
            ask2, node2, _ = e.cycle(s.n_asks)
            assert np.array_equal(ask2, want["ask"]) and np.array_equal(node2, want["node"]), commit


def test_user_limits(oracle):
    """user / group resource limits through the C ABI (yk_apps_user, yk_user_limits_set), both commits; the held amounts
    follow the cycle and yk_release"""
    for seed in (1, 5, 8, 13, 21):
        _check(synth.with_user_limits(synth.fuzz(seed), seed=seed), oracle, batch=64)
    s = synth.with_user_limits(synth.hier(300, 3, 4, 2, 40, seed=21), n_users=4, seed=1, frac=0.3)
    st = _check(s, oracle, batch=128)
    assert st["skipped"] > 0
    # second cycle after releasing everything: the limits bind again exactly as in the first
    want = oracle.run(s)
    with Engine.for_snapshot(s, batch=128) as e:
        ask, node, _ = e.cycle(s.n_asks)
        e.release(ask)
        e.load_snapshot(s)
        ask2, node2, _ = e.cycle(s.n_asks)
    assert np.array_equal(ask2, want["ask"]) and np.array_equal(node2, want["node"])


def test_reservation_phase_predicates(oracle):
    """Predicates(Allocate = false): same plugins without NodeResourcesFit / available (predicate_manager.go:353-368);
    device answer vs the oracle's, and the reference's reserve tables (TestReserveAlloc / TestReserveNodeSelector: the
    golden file's reserve cases fit in the reservation phase exactly when the test expects no error)"""
    s = synth.perf(60, 4, 30, masks=True, seed=77)
    s.node_avail[::3, 0] = 0                                       # full nodes: allocate phase fails, reserve phase may pass
    s.node_flags[5] = 0
    s.ask_node[7] = 3
    with Engine.for_snapshot(s) as e:
        diff = 0
        for a in range(0, s.n_asks, 5):
            for n in range(0, s.n_nodes, 3):
                got, exp = e.evaluate_reserve(a, n), oracle.predicate_reserve(s, a, n)
                assert got == exp, (a, n, got, exp)
                diff += int((got == 0) != (e.evaluate(a, n) == 0))
        assert diff > 0


def test_queues_set_reconciles_with_present_applications(oracle):
    """yk_queues_set while applications exist: a tree that drops or re-parents their queue is refused; with
    allocated == NULL what the applications hold is summed up the new tree, so quotas keep binding"""
    from yunikorn_k8shim_b200 import YkError
    s = synth.hier(60, 2, 2, 2, 20, seed=5)
    with Engine.for_snapshot(s, batch=32) as e:
        with pytest.raises(YkError) as ei:
            e.queues_set(s.q_parent[:3], s.q_guaranteed[:3], s.q_max[:3], s.q_alloc[:3], s.q_sort[:3])   # leaves are gone
        assert ei.value.code == -4
        par = s.q_parent.copy()
        par[-1] = s.n_queues - 2                                   # the last leaf's sibling becomes its parent: it holds apps
        with pytest.raises(YkError):
            e.queues_set(par, s.q_guaranteed, s.q_max, s.q_alloc, s.q_sort)
        ask, node, _ = e.cycle(40)                                  # some allocations
        want = oracle.run(s)
        assert np.array_equal(ask, want["ask"][:40])
        e.queues_set(s.q_parent, s.q_guaranteed, s.q_max, None, s.q_sort)   # same tree, allocated left to the library
        ask2, node2, _ = e.cycle(s.n_asks)
        assert np.array_equal(np.concatenate([ask, ask2]), want["ask"])     # quotas bound exactly as in one uncut cycle
        assert np.array_equal(np.concatenate([node, node2]), want["node"])
