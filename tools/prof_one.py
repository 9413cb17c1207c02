"""Runs the score call a few times on 4096 candidates of ONE shape type (or the c3 mix) -- the target of
an `ncu -k regex:score_kernel` capture.  usage: prof_one.py PLANE|SPHERE|CYLINDER|CONE|MIX [points] [reps]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ransac_jl_b200 as R
from ransac_jl_b200 import scenes
from ransac_jl_b200._lib import lib

which = sys.argv[1] if len(sys.argv) > 1 else "MIX"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 4 << 20
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
TYPES = ["PLANE", "SPHERE", "CYLINDER", "CONE"]
sc = scenes.scene_mixed(3, N)
if which == "MIX":
    cands = scenes.perturbed_candidates(sc, 1024, seed=7)
else:
    i = TYPES.index(which)
    cands = scenes.perturbed_candidates(sc, 4096, seed=7)[i * 4096:(i + 1) * 4096]
params = R.ransacparameters()
cp = R.to_c(params)
pc = R.RANSACCloud(sc.vertices, sc.normals, [np.zeros(0, np.int64)], device=0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
arr = R.pack_cands(cands)
d_cands = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
d_counts = torch.zeros(len(cands), dtype=torch.int32, device=dev)
for _ in range(reps):
    pc.ctx.check(lib.rsc_score_dev(pc.handle, C.byref(cp), d_cands.data_ptr(), len(cands), -1, d_counts.data_ptr(), stream.cuda_stream))
    torch.cuda.synchronize()
    ms, gp = C.c_double(), C.c_int64()
    pc.ctx.check(lib.rsc_ctx_last_kernel(pc.ctx.h, C.byref(ms), C.byref(gp)))
    print(which, N, "kernel ms", ms.value, "G evals/s", len(cands) * N / ms.value / 1e6)
