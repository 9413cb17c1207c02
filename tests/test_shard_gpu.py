"""Multi-rank runs of the device loop on ONE GPU: two processes share cuda:0 and sum their buffers
through a host-staged gloo all-reduce (rsc_ctx_set_allreduce) -- NCCL refuses two ranks on one device,
the library's own NCCL path is exercised by bench.py / tools/ransac_multi.py on multi-GPU boxes.

  * sharded storage (rsc_cloud_create_shard): every rank holds half of the cloud, minimal sets come from
    the replicated enabled mask, coordinates are gathered from their owners, inlier lists stay
    distributed -- the joined result equals the C oracle's loop (and so the single-GPU loop);
  * replicated storage with a cloud smaller than 2048 x ranks: one rank's range is empty and it still
    joins every all-reduce."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _scene(which):
    from ransac_jl_b200 import scenes

    if which == "c1":
        return scenes.scene_c1(), 2, {}, 4321
    if which == "tiny":
        sc = scenes.scene_c1()
        keep = np.random.default_rng(5).permutation(len(sc.vertices))[:1500]
        return scenes.Scene(sc.vertices[keep], sc.normals[keep], None, sc.primitives), 2, {"tau": 50, "itermax": 60}, 11
    return (scenes.scene_mixed(81, 40_000, noise_frac=0.002, jitter_deg=1.0, outlier_frac=0.1, counts=(3, 1, 1, 1)), 4,
            {"tau": 400, "minsubsetN": 128, "itermax": 120}, 99)


def _worker(rank, world, port, which, layout, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    dist.init_process_group("gloo", rank=rank, world_size=world)
    import ransac_jl_b200 as R
    from ransac_jl_b200 import shard

    sc, r, itp, seed = _scene(which)
    n = len(sc.vertices)
    params = R.ransacparameters(iteration=itp)
    subsets = R.makesubsets(n, r, np.random.default_rng(1234))
    if layout == "shard":
        lo, hi = shard.partition(n, world)[rank]
        pc = shard.ShardedCloud(sc.vertices[lo:hi], sc.normals[lo:hi], subsets, (lo, hi), n, comm="host")
        ex, _ = R.ransac(pc, params, True, seed=seed)
        assert all(len(e.inpoints) == 0 or (e.inpoints.min() >= lo and e.inpoints.max() < hi) for e in ex)
        local_en = pc.isenabled
        totals = [e.total for e in ex]
        ex = shard.gather_extracted(ex)
        parts = [None] * world
        dist.all_gather_object(parts, local_en)
        enabled = np.concatenate(parts)
    else:
        pc = R.RANSACCloud(sc.vertices, sc.normals, subsets)
        sh = shard.ShardedContext(pc, comm="host")
        ex, _ = R.ransac(pc, params, True, seed=seed)
        enabled = pc.isenabled
        totals = [e.total for e in ex]
        sh.close()
    if rank == 0:
        q.put(([(e.shape.to_cand().type, bool(e.shape.to_cand().outwards), list(e.shape.to_cand().p), e.inpoints) for e in ex],
               enabled, totals, pc.last_run_syncs, pc.last_run_batches))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("which,layout", [("c1", "shard"), ("noisy", "shard"), ("tiny", "replicated"), ("noisy", "replicated")])
def test_two_ranks_one_gpu_match_c_oracle(which, layout):
    from oracle import c_oracle
    import ransac_jl_b200 as R
    from tests.helpers import oracle_params

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(rk, 2, port, which, layout, q)) for rk in range(2)]
    for p in procs:
        p.start()
    import queue as _queue

    res = None
    for _ in range(120):  # a crashed worker must not leave the test waiting for the queue
        try:
            res = q.get(timeout=5)
            break
        except _queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert res is not None
    got, enabled, totals, syncs, batches = res
    sc, r, itp, seed = _scene(which)
    subsets = R.makesubsets(len(sc.vertices), r, np.random.default_rng(1234))
    want, en, info = c_oracle.ransac(sc.vertices, sc.normals, subsets[0], oracle_params(R.ransacparameters(iteration=itp)), seed)
    assert len(got) == len(want) and len(want) >= 2
    for g, w, tot in zip(got, want, totals):
        assert g[0] == w[0] and (g[0] == 0 or g[1] == w[1])
        np.testing.assert_allclose(g[2], w[2], rtol=1e-5, atol=1e-5 * max(1.0, np.abs(w[2]).max()))
        np.testing.assert_array_equal(g[3], w[3])
        assert tot == len(w[3])
    np.testing.assert_array_equal(enabled, en)
    # 2 synchronisations per batch + 1 per extraction (+ set-up)
    assert syncs <= 2 * batches + len(want) + 8


def test_library_nccl_communicator_single_rank():
    """the library's own NCCL path on one GPU: rsc_comm_unique_id + rsc_ctx_comm_init with one rank (libnccl.so.2
    is dlopen-ed here), then the device loop with a point range -- every count / hit / mask exchange goes through
    ncclAllReduce on the context stream and the result equals the plain run"""
    import ctypes as C

    import ransac_jl_b200 as R
    from ransac_jl_b200 import scenes
    from ransac_jl_b200._lib import lib

    sc, r, itp, seed = _scene("noisy")
    params = R.ransacparameters(iteration=itp)
    ctx = R.Context(0)  # a context of its own: the communicator stays with it
    pc = R.RANSACCloud(sc.vertices, sc.normals, r, ctx=ctx)
    plain, _ = R.ransac(pc, params, True, seed=seed)
    ident = (C.c_uint8 * 128)()
    ctx.check(lib.rsc_comm_unique_id(ident))
    ctx.check(lib.rsc_ctx_comm_init(ctx.h, ident, 0, 1))
    try:
        ctx.check(lib.rsc_cloud_set_range(pc.handle, 0, pc.size))
        viacomm, _ = R.ransac(pc, params, True, seed=seed)
        calls, nbytes = C.c_int64(), C.c_int64()
        ctx.check(lib.rsc_ctx_comm_stats(ctx.h, C.byref(calls), C.byref(nbytes)))
    finally:
        ctx.check(lib.rsc_ctx_comm_destroy(ctx.h))
    assert calls.value >= 2 * len(plain) and nbytes.value > 0
    assert len(plain) == len(viacomm) >= 3
    for a, b in zip(plain, viacomm):
        assert list(a.shape.to_cand().p) == list(b.shape.to_cand().p)
        np.testing.assert_array_equal(a.inpoints, b.inpoints)
    pc.close()
