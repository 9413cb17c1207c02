"""CPU tests of the multi-GPU host logic: the point-range partition, and -- with two gloo ranks --
that summing per-shard counts / disjoint inlier-mask words reproduces the unsharded result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_covers_exactly_once():
    from ransac_jl_b200.shard import ALIGN, partition

    for n in (1, 2047, 2048, 2049, 10_000, 1 << 20, 10_000_000, 100_000_001):
        for world in (1, 2, 4, 8):
            parts = partition(n, world)
            assert len(parts) == world
            assert parts[0][0] == 0 and parts[-1][1] == n
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c and a <= b
            for lo, hi in parts:
                assert lo % ALIGN == 0 and (hi % ALIGN == 0 or hi == n)


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import ransac_oracle as O
    from ransac_jl_b200 import scenes
    from ransac_jl_b200.shard import partition
    from tests.helpers import to_oracle_shape

    sc = scenes.scene_mixed(5, n)
    cands = scenes.perturbed_candidates(sc, 3, seed=2)
    lo, hi = partition(n, world)[rank]
    P, N = sc.vertices.astype(np.float64), sc.normals.astype(np.float64)
    op = O.default_parameters()
    # what each rank's GPU computes on its shard: counts and its slice of the inlier mask
    local = np.array([int(O.compatibles(to_oracle_shape(c), P[lo:hi], N[lo:hi], op).sum()) for c in cands], np.int32)
    words = np.zeros((n + 31) // 32, np.int32)
    m = np.zeros(n, bool)
    m[lo:hi] = O.compatibles(to_oracle_shape(cands[0]), P[lo:hi], N[lo:hi], op)
    packed = np.packbits(m, bitorder="little")
    words.view(np.uint8)[: len(packed)] = packed
    t_counts, t_words = torch.from_numpy(local), torch.from_numpy(words)
    dist.all_reduce(t_counts)  # C1: per-candidate counts
    dist.all_reduce(t_words)  # disjoint ranges: sum == union of the mask words
    if rank == 0:
        whole = np.array([int(O.compatibles(to_oracle_shape(c), P, N, op).sum()) for c in cands], np.int32)
        mask = O.compatibles(to_oracle_shape(cands[0]), P, N, op)
        got = np.unpackbits(t_words.numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
        q.put((bool(np.array_equal(t_counts.numpy(), whole)), bool(np.array_equal(got, mask))))
    dist.destroy_process_group()


def test_two_rank_gloo_count_and_mask_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 9000, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok_counts, ok_mask = q.get(timeout=5)
    assert ok_counts and ok_mask


def _worker_lists(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ransac_jl_b200 as R
    from ransac_jl_b200 import shard
    from tools import ransac_e2e

    n = 50_000
    rng = np.random.default_rng(9)  # the same "whole" result on every rank
    shapes = [R.FittedPlane(np.zeros(3), np.array([0.0, 0.0, 1.0])), R.FittedSphere(np.ones(3), 2.0, True)]
    whole = [np.sort(rng.choice(n, k, replace=False)).astype(np.int64) for k in (7000, 1200)]
    lo, hi = shard.partition(n, world)[rank]
    mine = [R.ExtractedShape(s, w[(w >= lo) & (w < hi)]) for s, w in zip(shapes, whole)]

    def allsum(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    dg = ransac_e2e.digest(mine, allsum)                     # additive over the ranks' parts
    full = ransac_e2e.digest([R.ExtractedShape(s, w) for s, w in zip(shapes, whole)], lambda a: a)
    joined = shard.gather_extracted(mine)                    # parts concatenated in rank order = the ascending list
    ok = dg == full and all(np.array_equal(j.inpoints, w) for j, w in zip(joined, whole))
    if rank == 0:
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_distributed_inlier_lists_digest_and_gather():
    """sharded storage leaves every inlier list distributed over the ranks: the digest bench.py compares across
    GPU counts is additive over the parts, and gather_extracted restores the ascending global lists"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_lists, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=5)


def _worker_comm_id(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ctypes as C

    from ransac_jl_b200 import shard
    from ransac_jl_b200._lib import lib

    ident = shard.exchange_comm_id()  # rank 0: the LIBRARY's rsc_comm_unique_id (dlopen of libnccl.so.2, no GPU needed)
    got = [None] * world
    dist.all_gather_object(got, ident)
    # without a context the collective initialisation must refuse (no NCCL call, no hang)
    buf = (C.c_uint8 * 128).from_buffer_copy(ident)
    rc_null = lib.rsc_ctx_comm_init(None, buf, rank, world)
    if rank == 0:
        q.put((len(ident), all(g == ident for g in got), any(ident), int(rc_null)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_library_nccl_id_reaches_every_rank():
    """the host plumbing of the in-library NCCL communicator (shard.init_comm = exchange_comm_id + rsc_ctx_comm_init):
    the 128-byte id drawn by the library on rank 0 arrives unchanged on every rank over a gloo group"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_comm_id, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    n, same, nonzero, rc_null = q.get(timeout=5)
    assert n == 128 and same and nonzero
    assert rc_null != 0
