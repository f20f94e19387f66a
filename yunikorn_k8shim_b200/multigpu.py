"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) for the exchange.

How the path shards (DESIGN.md "multi-GPU"): every rank holds the same node / ask tables and
runs the same ordering; the (ask x node) SWEEP of each batch is split by rows -- rank r evaluates the asks
[r*rows_per, (r+1)*rows_per) of the batch against all nodes -- then ONE all-gather per batch makes every
rank's fit rows visible everywhere and each rank applies the same ordered commit, so node state stays
replicated and the bindings are bit-identical on every rank and to the single-GPU run.  The exchange moves
rows_per * (N/8) bytes per rank per batch (about 0.6 MB for 512 asks x 10k nodes): latency-bound.

The engine (C ABI) knows nothing about torch: it calls the yk_allgather_fn registered with yk_set_exchange.
This module provides that callback on top of torch.distributed, and a per-cycle agreement check (one
all-reduce of a bindings hash) that turns silent divergence of the replicas into an error.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch


def shard_rows(n_rows: int, world: int, rank: int):
    """(first_row, rows_per, padded_total) exactly as yk_engine.cu run_batch computes them."""
    rows_per = (n_rows + world - 1) // world
    return min(n_rows, rank * rows_per), rows_per, rows_per * world


class _CudaView:
    """__cuda_array_interface__ wrapper around a raw device pointer (no copy)."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3,
                                         "strides": None}


def _tensor_from_pointer(ptr: int, nbytes: int, cuda: bool) -> torch.Tensor:
    if cuda:
        return torch.as_tensor(_CudaView(ptr, nbytes), device="cuda")
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return torch.from_numpy(np.frombuffer(buf, dtype=np.uint8))


def make_allgather(dist, cuda: bool = True, group=None):
    """Returns fn(ctx, buf, row_bytes, first_row, n_rows, total_rows, stream) -> 0/!=0 for yk_set_exchange:
    in-place all-gather of equal row blocks; the caller's rows are already in place at first_row."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)

    def fn(ctx, buf, row_bytes, first_row, n_rows, total_rows, stream):
        try:
            assert n_rows * world == total_rows, (n_rows, world, total_rows)
            whole = _tensor_from_pointer(int(buf), int(total_rows) * int(row_bytes), cuda)
            mine = whole[rank * n_rows * row_bytes:(rank + 1) * n_rows * row_bytes]
            if cuda:
                ext = torch.cuda.ExternalStream(int(stream))
                with torch.cuda.stream(ext):
                    dist.all_gather_into_tensor(whole, mine, group=group)
            else:
                parts = [torch.empty_like(mine) for _ in range(world)]
                dist.all_gather(parts, mine.clone(), group=group)
                for r, p in enumerate(parts):
                    whole[r * n_rows * row_bytes:(r + 1) * n_rows * row_bytes].copy_(p)
            return 0
        except Exception as exc:  # never let an exception cross the C boundary
            import sys
            sys.stderr.write(f"[ykgpu exchange] {exc!r}\n")
            return 1
    return fn


def attach(engine, dist, group=None, p2p=None):
    """Wire an Engine created with rank/world to the other ranks.

    Preferred: peer-to-peer (all ranks on one NVLink/NVSwitch node): CUDA IPC handles of the fit buffers are swapped
    once through torch.distributed, after which the sweep kernel stores its rows straight into every rank's buffer and
    sequence flags in peer memory order the batches -- no collective launch per batch.  Fallback (p2p=False, or IPC not
    available): one NCCL all-gather per batch through the yk_set_exchange callback."""
    import os
    if p2p is None:
        p2p = os.environ.get("YK_NO_P2P") is None
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if p2p and 2 <= world <= 8:
        ok = 1
        try:
            mine = engine.peer_export()
        except Exception:
            mine, ok = b"", 0
        blobs = [None] * world
        dist.all_gather_object(blobs, (ok, mine), group=group)
        if all(b[0] for b in blobs):
            try:
                for r, (_, blob) in enumerate(blobs):
                    if r != rank:
                        engine.peer_import(r, blob)
                engine.peer_enable()
            except Exception:
                ok = 0
        else:
            ok = 0
        t = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)     # all or nobody
        if int(t.item()) == 1:
            return "p2p"
        if ok:
            raise RuntimeError("peer-to-peer exchange enabled on this rank but not on all ranks")
    engine.set_exchange(make_allgather(dist, cuda=True, group=group))
    return "nccl"


def fnv1a64(ask: np.ndarray, node: np.ndarray) -> int:
    """order-sensitive 64-bit digest of the bindings (vectorised polynomial hash; not the oracle's FNV)."""
    a = np.asarray(ask, dtype=np.uint64)
    n = np.asarray(node, dtype=np.uint64)
    with np.errstate(over="ignore"):
        i = np.arange(1, len(a) + 1, dtype=np.uint64)
        h = (a * np.uint64(0x9E3779B97F4A7C15) + n * np.uint64(0xC2B2AE3D27D4EB4F)) * (i * np.uint64(2) + np.uint64(1))
        return int(np.bitwise_xor.reduce(h)) if len(a) else 0


def check_agreement(dist, ask, node, device=None, group=None) -> bool:
    """One all-reduce per cycle: every replica must have produced the same bindings in the same order."""
    h = fnv1a64(ask, node)
    t = torch.tensor([h & 0x7FFFFFFFFFFFFFFF, -(h & 0x7FFFFFFFFFFFFFFF)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return int(t[0].item()) == -int(t[1].item())
