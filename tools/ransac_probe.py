"""Where the wall time of ransac() on the c4 scene goes: the C call, the device loop inside it, and the
fetch of the inlier lists."""
import ctypes as C
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ransac_jl_b200 as R
from ransac_jl_b200 import scenes, _lib
from ransac_jl_b200._lib import lib
from ransac_jl_b200.params import to_c

sc = scenes.scene_cad()
it = {"tau": len(sc.vertices) // 1000, "minsubsetN": 8192, "itermax": 400}
params = R.ransacparameters(iteration=it)
pc = R.RANSACCloud(sc.vertices, sc.normals, 32)
R.ransac(pc, R.ransacparameters(iteration=dict(it, itermax=2)), True, seed=1)
for rep in range(3):
    t0 = time.perf_counter()
    pc.enable_all()
    t1 = time.perf_counter()
    cp = to_c(params)
    run = C.c_void_p()
    pc.ctx.check(lib.rsc_ransac_run(pc.handle, C.byref(cp), 2024, C.byref(run)))
    t2 = time.perf_counter()
    tot = 0
    for i in range(lib.rsc_run_nshapes(run)):
        cand = _lib.rsc_cand(); n = C.c_int64()
        lib.rsc_run_shape(run, i, C.byref(cand), C.byref(n))
        idx = np.empty(n.value, dtype=np.int64)
        lib.rsc_run_inpoints(run, i, idx.ctypes.data)
        tot += n.value
    t3 = time.perf_counter()
    secs = lib.rsc_run_seconds(run)
    lib.rsc_run_destroy(run)
    t4 = time.perf_counter()
    print(f"rep {rep}: enable_all {1e3*(t1-t0):.1f} ms | rsc_ransac_run call {1e3*(t2-t1):.1f} ms (device loop inside {1e3*secs:.1f}) | "
          f"fetch {tot} indices {1e3*(t3-t2):.1f} ms | destroy {1e3*(t4-t3):.1f} ms | total {1e3*(t4-t0):.1f} ms")
