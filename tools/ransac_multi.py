"""Sharded end-to-end ransac() over the GPUs of one box (torchrun, one process per GPU).

Every rank builds the same seeded scene, runs the device loop with its point range + the NCCL
all-reduce callback, and rank 0 checks the result against an unsharded run on its own GPU.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29517 tools/ransac_multi.py --scene c2
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ransac_jl_b200 as R
from ransac_jl_b200 import scenes
from ransac_jl_b200.shard import ShardedContext

ap = argparse.ArgumentParser()
ap.add_argument("--scene", default="c2", choices=["c1", "c2", "c4", "c5small", "c5"])
ap.add_argument("--points", type=int, default=100_000_000, help="c5: cloud size")
ap.add_argument("--check", type=int, default=1)
ap.add_argument("--sampler", default="root", choices=["root", "octree"])
ap.add_argument("--levels", type=int, default=8, help="octree levels of the cell sampler")
ap.add_argument("--itermax", type=int, default=0)
ap.add_argument("--progressive", type=int, default=0, help="1: progressive subset scoring (RSC_SCORE_PROGRESSIVE)")
ap.add_argument("--lw-period", type=int, default=1, help="octree sampler: refresh the level weights every P iterations")
ap.add_argument("--lsq", type=int, default=0, help="1: least-squares refit before each extraction (RSC_REFIT_LSQ)")
args = ap.parse_args()

rank = int(os.environ.get("RANK", 0))
world = int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

if args.scene == "c1":
    sc, r, it = scenes.scene_c1(), 2, {}
elif args.scene == "c2":
    sc, r = scenes.scene_c2(), 32
    it = {"tau": sc.vertices.shape[0] // 100, "minsubsetN": 4096, "itermax": 200}
elif args.scene == "c4":
    sc, r = scenes.scene_cad(), 32
    it = {"tau": sc.vertices.shape[0] // 1000, "minsubsetN": 8192, "itermax": 400}
elif args.scene == "c5small":
    sc, r = scenes.scene_lidar(10_000_000), 64
    it = {"tau": sc.vertices.shape[0] // 500, "minsubsetN": 8192, "itermax": 200}
else:
    # c5: the ranks generate the 12.5 M-point chunks of the scene side by side into /dev/shm, then
    # every rank maps all of them (the cloud is replicated; scoring/refit are sharded by point range)
    n, ch = args.points, scenes.LIDAR_CHUNK
    nch = (n + ch - 1) // ch
    tag = f"/dev/shm/rsc_c5_{n}"
    for i in range(rank, nch, world):
        part = scenes.scene_lidar_chunk(i, min(ch, n - i * ch))
        np.save(f"{tag}_v{i}.npy", part.vertices)
        np.save(f"{tag}_n{i}.npy", part.normals)
    if world > 1:
        dist.barrier()
    V = np.concatenate([np.load(f"{tag}_v{i}.npy", mmap_mode="r") for i in range(nch)])
    Nn = np.concatenate([np.load(f"{tag}_n{i}.npy", mmap_mode="r") for i in range(nch)])
    if world > 1:
        dist.barrier()
    if rank == 0:
        for i in range(nch):
            os.remove(f"{tag}_v{i}.npy"), os.remove(f"{tag}_n{i}.npy")
    sc, r = scenes.Scene(V, Nn, None, scenes.lidar_primitives()[0]), 64
    it = {"tau": n // 500, "minsubsetN": 8192, "itermax": 200}
if args.itermax:
    it["itermax"] = args.itermax
params = R.ransacparameters(iteration=it)
pc = R.RANSACCloud(sc.vertices, sc.normals, r, device=local)
t_cells = None
if args.sampler == "octree":
    torch.cuda.synchronize()
    tc = time.perf_counter()
    pc.build_cells(args.levels)
    t_cells = time.perf_counter() - tc
sh = ShardedContext(pc) if world > 1 else None
R.ransac(pc, R.ransacparameters(iteration=dict(it, itermax=2)), True, seed=1, sampler=args.sampler)  # warm-up (allocations, NCCL)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
ex, secs = R.ransac(pc, params, True, seed=2024, sampler=args.sampler, lsq=bool(args.lsq), progressive=bool(args.progressive),
                    lw_period=args.lw_period)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
tt = torch.tensor([dt], device="cuda")
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
ok = None
if rank == 0 and args.check and world > 1:
    sh.close()
    pc2 = R.RANSACCloud(sc.vertices, sc.normals, pc.subsets, device=local)
    if args.sampler == "octree":
        pc2.build_cells(args.levels)
    ex2, _ = R.ransac(pc2, params, True, seed=2024, sampler=args.sampler, lsq=bool(args.lsq), progressive=bool(args.progressive))
    ok = len(ex) == len(ex2) and all(
        list(a.shape.to_cand().p) == list(b.shape.to_cand().p) and np.array_equal(a.inpoints, b.inpoints) for a, b in zip(ex, ex2))
if rank == 0:
    st = pc.ctx.stats()
    print(json.dumps({"scene": args.scene, "points": int(sc.vertices.shape[0]), "n_gpus": world, "ransac_seconds": float(tt.item()),
                      "shapes": [[R.strt(e.shape), int(len(e.inpoints))] for e in ex][:40], "n_shapes": len(ex),
                      "points_extracted": int(sum(len(e.inpoints) for e in ex)), "evals": int(st.evals),
                      "sets_drawn": int(st.sets_drawn), "matches_unsharded": ok, "sampler": args.sampler, "lw_period": args.lw_period, "lsq": bool(args.lsq), "progressive": bool(args.progressive), "refined": int(getattr(pc, "last_refined", 0)),
                      "device_loop_seconds": float(getattr(pc, "last_run_seconds", float("nan"))),
                      "build_cells_seconds": t_cells,
                      "levelweight": [round(float(x), 4) for x in getattr(pc, "levelweight", [])]}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
