"""Cost of the first culled call on a subset (it builds the Morton view: sort + gather + spheres) against later calls,
on fresh clouds of 20 M points with 1.5 M-point subsets (the c5 subset size).   python tools/cull_view_bench.py"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ransac_jl_b200 as R
from ransac_jl_b200 import scenes
sc = scenes.scene_c3(20_000_000)
params = R.ransacparameters()
cands = scenes.perturbed_candidates(sc, 75, seed=7)
for rep in range(4):
    pc = R.RANSACCloud(sc.vertices, sc.normals, 13)
    pc.count_enabled()
    R.score_counts(pc, cands, 0, params)
    t0 = time.perf_counter()
    got, info = R.score_counts_culled(pc, cands, params, subsetID=0)
    t1 = time.perf_counter()
    got, info = R.score_counts_culled(pc, cands, params, subsetID=0)
    t2 = time.perf_counter()
    print(rep, len(pc.subsets[0]), "first culled call %.1f ms, second %.1f ms, kernel %.3f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t1), info["kernel_ms"]), flush=True)
    pc.close()
