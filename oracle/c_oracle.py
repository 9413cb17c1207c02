"""ctypes wrapper of oracle/liboracle.so (the C restatement of the reference).  Test infrastructure:
only tests/, smoke() and bench.py's cpu_baseline / --impl reference legs may import this."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


class orc_cand(C.Structure):
    _fields_ = [("type", C.c_int32), ("outwards", C.c_int32), ("p", C.c_double * 7)]


class orc_params(C.Structure):
    _fields_ = [
        ("eps", C.c_double * 4),
        ("alpha", C.c_double * 4),
        ("parallelthrdeg", C.c_double),
        ("sphere_par", C.c_double),
        ("minconeopang", C.c_double),
        ("collin_threshold", C.c_double),
        ("shape_types", C.c_int32 * 4),
        ("n_shape_types", C.c_int32),
        ("sphere_ignores_enabled", C.c_int32),
    ]


def available() -> bool:
    return _load() is not None


def _load():
    global _lib
    if _lib is None and os.path.exists(_PATH):
        _lib = C.CDLL(_PATH)
        _lib.orc_score.restype = C.c_int
        _lib.orc_fit_points.restype = C.c_int
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def to_params(op: dict, sphere_ignores_enabled: bool = True) -> orc_params:
    """oracle-style params dict (shape_types are kinds) -> orc_params"""
    import math

    p = orc_params()
    for name, k in (("plane", 0), ("sphere", 1), ("cylinder", 2), ("cone", 3)):
        g = op.get(name, {"eps": 0.3, "alpha": math.radians(5)})
        p.eps[k], p.alpha[k] = g["eps"], g["alpha"]
    p.parallelthrdeg = op["common"]["parallelthrdeg"]
    p.collin_threshold = op["common"]["collin_threshold"]
    p.sphere_par = op.get("sphere", {}).get("sphere_par", 0.02)
    p.minconeopang = op.get("cone", {}).get("minconeopang", math.radians(2))
    st = op["iteration"]["shape_types"]
    p.n_shape_types = len(st)
    for i, t in enumerate(st):
        p.shape_types[i] = int(t)
    p.sphere_ignores_enabled = int(sphere_ignores_enabled)
    return p


def pack(shapes) -> "C.Array":
    """shapes: objects with .to_cand() (package host types) or oracle Shapes"""
    arr = (orc_cand * max(1, len(shapes)))()
    for i, s in enumerate(shapes):
        if hasattr(s, "to_cand"):
            c = s.to_cand()
            arr[i].type, arr[i].outwards = c.type, c.outwards
            arr[i].p[:] = list(c.p)
        else:
            arr[i].type, arr[i].outwards = s.kind, int(s.outwards)
            arr[i].p[:] = list(s.params7())
    return arr


def score_counts(shapes, P, N, op, enabled=None, want_masks=False, nthreads=0):
    """(counts, threads_used[, masks]) -- compatibles* + count over points P, N (n x 3 float64)."""
    lib = _load()
    P = np.ascontiguousarray(P, dtype=np.float64)
    N = np.ascontiguousarray(N, dtype=np.float64)
    n = len(P)
    arr = pack(shapes)
    prm = to_params(op)
    counts = np.zeros(len(shapes), np.int32)
    masks = np.zeros((len(shapes), n), np.uint8) if want_masks else None
    en = None if enabled is None else np.ascontiguousarray(enabled, dtype=np.uint8)
    used = lib.orc_score(arr, len(shapes), P.ctypes.data_as(C.c_void_p), N.ctypes.data_as(C.c_void_p), C.c_int64(n),
                         en.ctypes.data_as(C.c_void_p) if en is not None else None, C.byref(prm),
                         counts.ctypes.data_as(C.c_void_p), masks.ctypes.data_as(C.c_void_p) if want_masks else None,
                         C.c_int(nthreads))
    return (counts, used, masks.astype(bool)) if want_masks else (counts, used)


def fit_points(P, N, op):
    """forcefitshapes! over S sets: P, N = (S, k, 3).  Returns (list of (type, outwards, p[7]), out_set)."""
    lib = _load()
    P = np.ascontiguousarray(P, dtype=np.float64)
    N = np.ascontiguousarray(N, dtype=np.float64)
    S, k = P.shape[0], P.shape[1]
    prm = to_params(op)
    out = (orc_cand * max(1, S * prm.n_shape_types))()
    out_set = np.zeros(max(1, S * prm.n_shape_types), np.int32)
    m = lib.orc_fit_points(P.ctypes.data_as(C.c_void_p), N.ctypes.data_as(C.c_void_p), C.c_int(S), C.c_int(k), C.byref(prm), out,
                           out_set.ctypes.data_as(C.c_void_p))
    return [(out[i].type, bool(out[i].outwards), np.array(out[i].p[:])) for i in range(m)], out_set[:m].copy()


class orc_iter(C.Structure):
    _fields_ = [("drawN", C.c_int32), ("minsubsetN", C.c_int32), ("itermax", C.c_int32), ("extract_s", C.c_int32),
                ("terminate_s", C.c_int32), ("reserved", C.c_int32), ("tau", C.c_int64), ("prob_det", C.c_double)]


_S = {"lengthC": 0, "allcand": 1, "nofminset": 2}


def estimate_score(s1len: int, plen: int, sigma: int):
    """estimatescore (confidenceintervals.jl:53-74) with Int64 wrap-around -> (min, max, E)"""
    lib = _load()
    a, b, e = C.c_double(), C.c_double(), C.c_double()
    lib.orc_estimate_score(C.c_int64(s1len), C.c_int64(plen), C.c_int64(sigma), C.byref(a), C.byref(b), C.byref(e))
    return a.value, b.value, e.value


def ransac(P, N, subset1, op, seed, enabled=None, nthreads=0):
    """The whole loop (iterations.jl:35-162) in C on the Philox minimal sets of the NumPy oracle.

    P, N: (n, 3) float64; subset1: pc.subsets[0] (0-based global indices); enabled: bool (n,) or None
    (= ransac(pc, params, true)).  Returns (shapes, enabled_after, info): shapes = list of
    (type, outwards, p[7], inpoints int64 ascending), info = dict(iterations, cands_scored, evals,
    extracted_at, seconds=(sample+fit, score, refit+bookkeeping), threads)."""
    lib = _load()
    lib.orc_ransac.restype = C.c_void_p
    lib.orc_run_shape.restype = C.c_int64
    lib.orc_run_cands_scored.restype = C.c_int64
    lib.orc_run_evals.restype = C.c_int64
    P = np.ascontiguousarray(P, dtype=np.float64)
    N = np.ascontiguousarray(N, dtype=np.float64)
    sub = np.ascontiguousarray(subset1, dtype=np.int64)
    n = len(P)
    en = np.ones(n, np.uint8) if enabled is None else np.ascontiguousarray(enabled, dtype=np.uint8).copy()
    prm = to_params(op)
    i = op["iteration"]
    it = orc_iter(int(i["drawN"]), int(i["minsubsetN"]), int(i["itermax"]), _S[i["extract_s"]], _S[i["terminate_s"]], 0,
                  int(i["tau"]), float(i["prob_det"]))
    run = C.c_void_p(lib.orc_ransac(P.ctypes.data_as(C.c_void_p), N.ctypes.data_as(C.c_void_p), C.c_int64(n),
                                    sub.ctypes.data_as(C.c_void_p), C.c_int64(len(sub)), en.ctypes.data_as(C.c_void_p),
                                    C.byref(prm), C.byref(it), C.c_uint64(seed), C.c_int(nthreads)))
    try:
        shapes, at = [], []
        for k in range(lib.orc_run_nshapes(run)):
            c, a = orc_cand(), C.c_int32()
            ln = lib.orc_run_shape(run, k, C.byref(c), C.byref(a))
            idx = np.empty(ln, np.int64)
            if ln:
                lib.orc_run_inpoints(run, k, idx.ctypes.data_as(C.c_void_p))
            shapes.append((c.type, bool(c.outwards), np.array(c.p[:]), idx))
            at.append(a.value)
        secs = (C.c_double * 3)()
        lib.orc_run_seconds(run, secs)
        info = {"iterations": lib.orc_run_iterations(run), "cands_scored": lib.orc_run_cands_scored(run),
                "evals": lib.orc_run_evals(run), "extracted_at": at, "seconds": tuple(secs),
                "threads": nthreads if nthreads > 0 else lib.orc_max_threads()}
    finally:
        lib.orc_run_free(run)
    return shapes, en.astype(bool), info
