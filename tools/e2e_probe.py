"""Times the pieces of the host-buffer (e2e) step separately: cloud upload, score, destroy."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ransac_jl_b200 as R
from ransac_jl_b200 import scenes
from ransac_jl_b200._lib import lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16 << 20
sc = scenes.scene_mixed(3, n)
cands = scenes.perturbed_candidates(sc, 1024, seed=7)
params = R.ransacparameters()
cp = R.to_c(params)
arr = R.pack_cands(cands)
ctx = R.Context.get(0)
xyz = torch.from_numpy(sc.vertices).pin_memory()
nrm = torch.from_numpy(sc.normals).pin_memory()
counts = np.zeros(len(cands), np.int32)
for it in range(4):
    t0 = time.perf_counter()
    h = C.c_void_p()
    ctx.check(lib.rsc_cloud_create(ctx.h, xyz.data_ptr(), nrm.data_ptr(), n, C.byref(h)))
    t1 = time.perf_counter()
    ctx.check(lib.rsc_score(h, C.byref(cp), arr, len(cands), -1, counts.ctypes.data, None))
    t2 = time.perf_counter()
    lib.rsc_cloud_destroy(h)
    t3 = time.perf_counter()
    st = ctx.stats()
    print(f"iter {it}: create {1e3*(t1-t0):.1f} ms, score {1e3*(t2-t1):.1f} ms (events: call {st.score_ms:.1f}, kernel {st.last_kernel_ms:.1f}), destroy {1e3*(t3-t2):.1f} ms, guard tasks {st.exact_pairs}")
