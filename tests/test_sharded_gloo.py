"""CPU, world_size 2, gloo: the exchange plumbing of the feature-sharded fit (sharded.py).

The kernels need a GPU, but the two collectives of a level do not: the candidate all-gather must
deliver the [rank][node] layout b200dt_session_decide reads, and the sum-all-reduce of the left-row
masks must equal their OR (bits of different nodes are disjoint).  The host-side winner rule is
checked against the oracle's tie rule on the gathered records.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from np_decision_tree_b200 import _lib
from np_decision_tree_b200.sharded import gather_candidates, reduce_side_mask, shard_columns

N_NODES = 5
N_ROWS = 1000


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_candidates(rank, world):
    """Deterministic per-rank candidate records with cross-rank gain ties."""
    rec = np.zeros(N_NODES, dtype=[("gain", "f8"), ("aux", "f8"), ("left_sum", "f8"), ("gain2", "f8"),
                                   ("threshold", "f8"), ("k", "i4"), ("f", "i4"), ("exact", "i4"), ("pad", "i4")])
    assert rec.itemsize == _lib.CANDIDATE_BYTES
    lo, hi = shard_columns(10, world, rank)
    for node in range(N_NODES):
        rec[node] = (float((node * 7 + rank * 3) % 5), 1.0 + rank, 0.5 * node, -1.0, 0.25 * rank,
                     (node + rank) % 3, lo + node % (hi - lo), 1, 0)
    return rec


def _pick(cands):
    """(gain desc, k asc, aux desc, f asc) -- the reference's tie rule, ref one.py:59,62."""
    return min(range(len(cands)), key=lambda i: (-cands[i]["gain"], cands[i]["k"], -cands[i]["aux"], cands[i]["f"]))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = _rank_candidates(rank, world)
        buf = torch.from_numpy(mine.view(np.uint8).copy())
        gathered = gather_candidates(buf, world).numpy().view(mine.dtype).reshape(world, N_NODES)
        for r in range(world):
            assert np.array_equal(gathered[r], _rank_candidates(r, world)), "layout must be [rank][node]"
        winners = [int(_pick([gathered[r][node] for r in range(world)])) for node in range(N_NODES)]
        # each winner's owner marks that node's left rows; nodes own disjoint row ranges
        mask = torch.zeros((N_ROWS + 31) // 32, dtype=torch.int32)
        expect = np.zeros((N_ROWS + 31) // 32, dtype=np.uint32)
        per = N_ROWS // N_NODES
        for node, w in enumerate(winners):
            rows = np.arange(node * per, node * per + per // 2)
            bits = np.zeros_like(expect)
            np.bitwise_or.at(bits, rows >> 5, np.uint32(1) << (rows & 31).astype(np.uint32))
            expect |= bits
            if w == rank:
                mask |= torch.from_numpy(bits.view(np.int32))
        reduce_side_mask(mask, world)
        assert np.array_equal(mask.numpy().view(np.uint32), expect), "SUM of disjoint masks must equal OR"
        # the totals round of the sharded exact search (include/b200dt.h, b200dt_session_begin): every rank emits a
        # record marked exact == 2; only the owner of the last global column (f == F - 1) carries the node total
        F = 10
        lo, hi = shard_columns(F, world, rank)
        owner = hi == F
        tot = np.zeros(N_NODES, dtype=mine.dtype)
        for node in range(N_NODES):
            tot[node] = (0.0, 0.0, 100.0 + node if owner else 0.0, 0.0, 0.0, -1, F - 1 if owner else -1, 2, 0)
        g2 = gather_candidates(torch.from_numpy(tot.view(np.uint8).copy()), world).numpy().view(mine.dtype)
        g2 = g2.reshape(world, N_NODES)
        totals = []
        for node in range(N_NODES):
            recs = [g2[r][node] for r in range(world)]
            assert all(int(c["exact"]) == 2 for c in recs)
            have = [float(c["left_sum"]) for c in recs if int(c["f"]) == F - 1]
            assert len(have) == 1, "exactly one rank owns the last global column"
            totals.append(have[0])
        assert totals == [100.0 + node for node in range(N_NODES)]
        out[rank] = winners
    finally:
        dist.destroy_process_group()


def test_exchange_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0] == out[1], "every rank must reach the same decision from the same gathered records"
    assert len(out[0]) == N_NODES


def test_world1_is_a_no_op():
    buf = torch.arange(56, dtype=torch.uint8)
    assert gather_candidates(buf, 1) is buf
    m = torch.ones(4, dtype=torch.int32)
    assert reduce_side_mask(m, 1) is m


# ---- classification sessions: 32-byte candidates, and the left class counts travel like the mask ----
def _class_candidates(rank, world, n_cols=9):
    rec = np.zeros(N_NODES, dtype=[("score", "f8"), ("threshold", "f8"), ("k", "i4"), ("f", "i4"), ("k_ext", "i4"),
                                   ("pad", "i4")])
    assert rec.itemsize == _lib.CLASS_CANDIDATE_BYTES
    lo, hi = shard_columns(n_cols, world, rank)
    for node in range(N_NODES):
        k = (node + rank) % 3
        rec[node] = (float((node * 5 + rank * 2) % 4), 0.5 * rank, k, lo + node % (hi - lo), k + node % 2, 0)
    return rec


def _class_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_classes = 3
        mine = _class_candidates(rank, world)
        buf = torch.from_numpy(mine.view(np.uint8).copy())
        gathered = gather_candidates(buf, world).numpy().view(mine.dtype).reshape(world, N_NODES)
        for r in range(world):
            assert np.array_equal(gathered[r], _class_candidates(r, world))
        # flat argmax over (position, GLOBAL column): score desc, k asc, f asc (ref strategy.py:277-278)
        winners = [min(range(world), key=lambda r: (-gathered[r][node]["score"], gathered[r][node]["k"],
                                                    gathered[r][node]["f"])) for node in range(N_NODES)]
        # only the owner of a winning column knows the left child's class counts; the others send zeros
        counts = torch.zeros(N_NODES * n_classes, dtype=torch.int32)
        expect = np.zeros((N_NODES, n_classes), dtype=np.int32)
        for node, w in enumerate(winners):
            expect[node] = [node + 1, 2 * node, gathered[w][node]["k_ext"]]
            if w == rank:
                counts.view(N_NODES, n_classes)[node] = torch.from_numpy(expect[node])
        reduce_side_mask(counts, world)
        assert np.array_equal(counts.numpy().reshape(N_NODES, n_classes), expect)
        out[rank] = winners
    finally:
        dist.destroy_process_group()


def test_class_exchange_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_class_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0] == out[1] and len(out[0]) == N_NODES
