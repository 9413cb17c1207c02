"""Dense vs Morton-tile-culled whole-cloud scoring on the c3 workload (one GPU).
    python tools/cull_bench.py --points 16777216 --cands 4096"""
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
ap.add_argument("--points", type=int, default=16 << 20)
ap.add_argument("--cands", type=int, default=4096)
ap.add_argument("--levels", type=int, default=11)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
sc = scenes.scene_c3(args.points)
cands = scenes.perturbed_candidates(sc, args.cands // 4, seed=7)
pc = R.RANSACCloud(sc.vertices, sc.normals, 1)
params = R.ransacparameters()
t0 = time.perf_counter()
pc.build_cells(args.levels)
pc.count_enabled()
t_cells = time.perf_counter() - t0
dense, _ = R.score_counts(pc, cands, -1, params)
t0 = time.perf_counter()
got, info = R.score_counts_culled(pc, cands, params)  # first call builds the Morton copy + tile spheres
t_first = time.perf_counter() - t0
ms = []
for _ in range(args.reps):
    got, info = R.score_counts_culled(pc, cands, params)
    ms.append(info["kernel_ms"])
dms = []
for _ in range(args.reps):
    R.score_counts(pc, cands, -1, params)
    dms.append(pc.ctx.stats().last_kernel_ms)
evals = len(cands) * pc.size
print(json.dumps({"points": pc.size, "candidates": len(cands), "equal_counts": bool(np.array_equal(got, dense)),
                  "mismatches": int((got != dense).sum()), "build_cells_s": round(t_cells, 4), "first_call_s": round(t_first, 4),
                  "culled_kernel_ms": ms, "dense_kernel_ms": dms, "pairs_total": info["pairs_total"],
                  "pairs_survived": info["pairs_survived"], "surviving_fraction": info["pairs_survived"] / info["pairs_total"],
                  "culled_G_evals_s_equivalent": evals / (min(ms) * 1e-3) / 1e9, "dense_G_evals_s": evals / (min(dms) * 1e-3) / 1e9}))
