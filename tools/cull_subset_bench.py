"""Culled vs dense scoring of a cell-sampler batch (tens of thousands of LOCAL candidates) against subset 1 of the
c4 scene -- the K2 call of the octree loop.   python tools/cull_subset_bench.py --sets 65536"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ransac_jl_b200 as R
from ransac_jl_b200 import scenes

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=10_000_000)
ap.add_argument("--subsets", type=int, default=32)
ap.add_argument("--sets", type=int, default=65536)
ap.add_argument("--levels", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
sc = scenes.scene_cad(args.points)
pc = R.RANSACCloud(sc.vertices, sc.normals, args.subsets)
params = R.ransacparameters(iteration={"tau": args.points // 1000, "minsubsetN": 8192, "itermax": 400})
pc.build_cells(args.levels)
lw = np.full(args.levels, 1.0 / args.levels)
cands, _, _, _ = R.sample_fit_cells(pc, params, 1, 0, args.sets, lw)
dense, _ = R.score_counts(pc, cands, 0, params)
got, info = R.score_counts_culled(pc, cands, params, subsetID=0)
ms, dms = [], []
for _ in range(args.reps):
    got, info = R.score_counts_culled(pc, cands, params, subsetID=0)
    ms.append(info["kernel_ms"])
for _ in range(args.reps):
    t0 = time.perf_counter()
    R.score_counts(pc, cands, 0, params)
    dms.append(pc.ctx.stats().last_kernel_ms)
print(json.dumps({"subset_points": len(pc.subsets[0]), "candidates": len(cands), "equal_counts": bool(np.array_equal(got, dense)),
                  "culled_ms": ms, "dense_kernel_ms": dms, "surviving_fraction": info["pairs_survived"] / max(1, info["pairs_total"]),
                  "pairs_survived": info["pairs_survived"]}))
