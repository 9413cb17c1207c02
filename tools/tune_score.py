"""Tiling sweep of the K2 score kernel on one B200: per shape type, times every tiling in the
library's table (RSC_CFG_<TYPE>="K,MINB,U" hook) on 4096 candidates of that type x N points, then
the c3 mix with the per-type winners.  Kernel time = the library's CUDA events around the score
launches.  Prints one JSON line per measurement (-> profiles/)."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ransac_jl_b200 as R
from ransac_jl_b200 import scenes
from ransac_jl_b200._lib import lib

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8 << 20
CFGS = (os.environ.get("TUNE_CFGS") or "4,4,1;4,3,1;4,4,3;4,3,3;8,2,1;4,4,5;4,3,5;4,4,6;4,3,6;8,2,5").split(";")
TYPES = ["PLANE", "SPHERE", "CYLINDER", "CONE"]
FLOPS = {"PLANE": 13, "SPHERE": 16, "CYLINDER": 27, "CONE": 38}

sc = scenes.scene_mixed(3, N)
allc = scenes.perturbed_candidates(sc, 4096, seed=7)
by = {t: allc[i * 4096:(i + 1) * 4096] for i, t in enumerate(TYPES)}
mix = scenes.perturbed_candidates(sc, 1024, seed=7)
params = R.ransacparameters()
cp = R.to_c(params)
pc = R.RANSACCloud(sc.vertices, sc.normals, [np.zeros(0, np.int64)], device=0)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)


def time_set(cands, reps=3):
    arr = R.pack_cands(cands)
    d_cands = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
    d_counts = torch.zeros(len(cands), dtype=torch.int32, device=dev)
    best = None
    for _ in range(reps + 1):
        pc.ctx.check(lib.rsc_score_dev(pc.handle, C.byref(cp), d_cands.data_ptr(), len(cands), -1, d_counts.data_ptr(),
                                       stream.cuda_stream))
        torch.cuda.synchronize()
        ms, gp = C.c_double(), C.c_int64()
        pc.ctx.check(lib.rsc_ctx_last_kernel(pc.ctx.h, C.byref(ms), C.byref(gp)))
        best = ms.value if best is None else min(best, ms.value)
    return best, int(d_counts.sum().item()), gp.value


peak = 148 * 128 * 2 * 1.965e9 / 1e12
winners = {}
for t in TYPES:
    ref_sum = None
    for cfg in CFGS:
        os.environ["RSC_CFG_" + t] = cfg
        ms, csum, gp = time_set(by[t])
        if ref_sum is None:
            ref_sum = csum
        ge = 4096 * N / ms / 1e6
        print(json.dumps({"type": t, "cfg_K_minb_U": cfg, "kernel_ms": round(ms, 3), "G_evals_s": round(ge, 1),
                          "frac_fp32_peak": round(ge * FLOPS[t] / 1e3 / peak, 4), "counts_ok": csum == ref_sum,
                          "guard_pairs": gp}), flush=True)
        if t not in winners or ms < winners[t][1]:
            winners[t] = (cfg, ms)
    os.environ["RSC_CFG_" + t] = winners[t][0]
print(json.dumps({"winners": {t: winners[t][0] for t in TYPES}}), flush=True)
for waves in (16, 32, 64):
    for nofork in (False, True):
        os.environ["RSC_WAVES"] = str(waves)
        if nofork:
            os.environ["RSC_NOFORK"] = "1"
        else:
            os.environ.pop("RSC_NOFORK", None)
        ms, csum, gp = time_set(mix)
        ge = 4096 * N / ms / 1e6
        print(json.dumps({"type": "mix", "waves": waves, "fork": not nofork, "kernel_ms": round(ms, 3),
                          "G_evals_s": round(ge, 1), "frac_fp32_peak": round(ge * 23.5 / 1e3 / peak, 4)}), flush=True)
