"""N>1 host-side path on CPU: world_size-2 gloo run of the exchange callback the engine calls once per batch
(in-place all-gather of row shards) and of the per-cycle agreement check."""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from yunikorn_k8shim_b200 import multigpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_rows, row_bytes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, rows_per, total = multigpu.shard_rows(n_rows, world, rank)
        buf = np.zeros(total * row_bytes, dtype=np.uint8)
        view = buf.reshape(total, row_bytes)
        view[rank * rows_per:(rank + 1) * rows_per] = (np.arange(rows_per)[:, None] * 7 + rank * 101 + np.arange(row_bytes)[None, :]) % 251
        fn = multigpu.make_allgather(dist, cuda=False)
        rc = fn(None, buf.ctypes.data, row_bytes, rank * rows_per, rows_per, total, None)
        want = np.zeros_like(view)
        for r in range(world):
            want[r * rows_per:(r + 1) * rows_per] = (np.arange(rows_per)[:, None] * 7 + r * 101 + np.arange(row_bytes)[None, :]) % 251
        ok_gather = rc == 0 and np.array_equal(view, want)
        same = multigpu.check_agreement(dist, np.arange(10), np.arange(10) * 3)
        diff = multigpu.check_agreement(dist, np.arange(10), np.arange(10) * 3 + (rank == 1))
        q.put((rank, bool(ok_gather), bool(same), bool(diff), first, rows_per, total))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [64, 1001])
def test_exchange_and_agreement_world2(n_rows):
    world, row_bytes = 2, 40
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_rows, row_bytes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_gather, same, diff, first, rows_per, total in res:
        assert ok_gather, f"rank {rank}: gathered rows differ"
        assert same, "identical bindings must agree"
        assert not diff, "diverging replicas must be detected"
        assert rows_per == (n_rows + world - 1) // world and total == rows_per * world
        assert first == min(n_rows, rank * rows_per)


def test_shard_rows_covers_every_row_once():
    for n in (0, 1, 7, 64, 1000, 4096):
        for world in (1, 2, 4, 8):
            seen = np.zeros(n, dtype=int)
            for r in range(world):
                first, rows_per, total = multigpu.shard_rows(n, world, r)
                seen[first:min(n, first + rows_per)] += 1
                assert total >= n and total % world == 0
            assert (seen == 1).all()


# ---- the whole N>1 host path on CPU: two replicas of the engine's host side (tests/host/engine_shim.cpp: the real
# orderer and ordered commit, CPU stand-in for the sweep), rows of every batch cut across the ranks exactly as
# produce() cuts them and exchanged through the same callback type the engine calls (yk_allgather_fn) over gloo ----
def _replica(rank, world, port, seeds, split_min_pairs, q):
    import subprocess
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, here)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import test_engine_host as T
        from yunikorn_k8shim_b200 import synth
        from yunikorn_k8shim_b200.engine import ALLGATHER_FN
        from oracle import oracle_ctypes as oc
        so = os.path.join(os.environ["YK_SHIM_DIR"], f"engine_shim_{rank}.so")
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", so,
                               os.path.join(here, "host", "engine_shim.cpp")])
        shim = C.CDLL(so)
        cb = ALLGATHER_FN(multigpu.make_allgather(dist, cuda=False))
        out = []
        for seed in seeds:
            s = synth.fuzz(seed) if seed >= 0 else synth.perf(300, 12, 40, masks=(seed == -2))
            if (s.ask_gang >= 0).any() and np.bincount(s.ask_gang[s.ask_gang >= 0]).max() > 64:
                continue
            want = oc.run(s)
            real = shim.engine_host_run

            def ranked(*args):      # same argument list, prefixed with (rank, world, callback, threshold)
                return shim.engine_host_run_ranked(C.c_uint32(rank), C.c_uint32(world), cb, C.c_uint64(split_min_pairs), *args)

            class Proxy:
                engine_host_run = staticmethod(ranked)
            rc, ask, node, state, avail = T.run_engine_host(Proxy, s, batch=64)
            ok = rc == 0 and np.array_equal(ask, want["ask"]) and np.array_equal(node, want["node"]) and \
                np.array_equal(avail, want["avail"]) and np.array_equal(state, want["state"])
            agree = multigpu.check_agreement(dist, ask, node)
            out.append((seed, bool(ok), bool(agree)))
            assert real is not None
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("split_min_pairs", [0, 20000])
def test_two_replicas_split_rows_and_commit_identically(tmp_path, split_min_pairs):
    """split_min_pairs = 0: every batch is cut across the two ranks (rank 1 may own zero rows when rows are shared);
    20000: small sweeps stay whole on every rank and only the large ones are exchanged -- both must reproduce the
    oracle's bindings on both ranks, and the replicas must agree."""
    world = 2
    os.environ["YK_SHIM_DIR"] = str(tmp_path)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    seeds = [-1, -2, 0, 2, 5, 8, 13, 21, 34]
    procs = [ctx.Process(target=_replica, args=(r, world, port, seeds, split_min_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(res[0]) == len(res[1]) >= 6
    for rank in range(world):
        for seed, ok, agree in res[rank]:
            assert ok, f"rank {rank}, snapshot {seed}: differs from the oracle"
            assert agree, f"rank {rank}, snapshot {seed}: replicas disagree"
