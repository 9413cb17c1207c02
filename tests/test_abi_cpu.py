"""CPU-side checks of the boundary: the shared library loads, exports every symbol include/rsc.h
declares, the POD layouts match the header, and the host layer fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "rsc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rsc_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import ransac_jl_b200 as R

    syms = _declared_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(R._lib.lib, s), f"{s} declared in include/rsc.h but not exported"
        assert s in R._lib.SIGNATURES, f"{s} has no ctypes signature"
    assert R._lib.lib.rsc_version() == 100


def test_pod_layouts():
    import ransac_jl_b200 as R

    assert C.sizeof(R._lib.rsc_cand) == 64
    assert R._lib.rsc_cand.p.offset == 8
    assert C.sizeof(R._lib.rsc_params) == 160
    assert C.sizeof(R._lib.rsc_stats) == 64


def test_default_params_pod_matches_reference_defaults():
    # test/utilitytests.jl:41-82
    import math

    import ransac_jl_b200 as R

    p = R._lib.rsc_params()
    R._lib.lib.rsc_params_default(p)
    assert (p.drawN, p.minsubsetN, p.prob_det, p.tau, p.itermax) == (3, 15, 0.9, 900, 1000)
    assert list(p.shape_types) == [0, 3, 2, 1] and p.n_shape_types == 4  # plane, cone, cylinder, sphere
    assert list(p.eps) == [0.3] * 4
    assert all(abs(a - math.radians(5)) < 1e-16 for a in p.alpha)
    assert (p.sphere_par, p.collin_threshold, p.parallelthrdeg) == (0.02, 0.2, 1.0)
    assert abs(p.minconeopang - math.radians(2)) < 1e-16
    c = R.to_c(R.ransacparameters(sphere={"eps": 0.01}, iteration={"tau": 50}))
    assert c.eps[1] == 0.01 and c.eps[0] == 0.3 and c.tau == 50


def test_parameter_mirror_matches_reference_tests():
    # test/utilitytests.jl:84-114
    import math

    import ransac_jl_b200 as R

    p = R.ransacparameters([R.FittedSphere, R.FittedCylinder], sphere={"eps": 0.01}, cylinder={"alpha": 0.02})
    assert p["sphere"] == {"eps": 0.01, "alpha": math.radians(5), "sphere_par": 0.02}
    assert p["cylinder"] == {"eps": 0.3, "alpha": 0.02}
    assert p["iteration"]["shape_types"] == [R.FittedSphere, R.FittedCylinder]
    assert R.DEFAULT_PARAMETERS["iteration"]["shape_types"] == [R.FittedPlane, R.FittedCone, R.FittedCylinder, R.FittedSphere]


def test_confidence_interval_mirror():
    # test/confidenceintervals.jl:1-25
    import ransac_jl_b200 as R

    ci = R.ConfidenceInterval(1.0, 3)
    assert (ci.E, ci.min, ci.max) == (2.0, 1.0, 3.0)
    with pytest.raises(ValueError):
        R.ConfidenceInterval(3, 1.0)
    nc = R.notsoconfident(153.9, 9.7)
    assert (nc.min, nc.max, nc.E) == (9.7, 153.9, 81.8)


def test_iteration_candidates_store():
    # test/fitting.jl:1-18
    import ransac_jl_b200 as R

    ic = R.IterationCandidates()
    fp = R.FittedPlane([0.5, 0.5, 0.5], [0, 0, 1.0])
    assert len(ic) == 0
    ic.recordscore(fp, R.ConfidenceInterval(0, 1), [1, 2, 3, 4, 5])
    assert len(ic) == 1 and len(ic.scores) == 1 and len(ic.inpoints) == 1
    ic.deleteat(0)
    assert len(ic) == 0 and len(ic.shapes) == 0


def test_estimatescore_through_abi_matches_oracle():
    import ransac_jl_b200 as R
    from oracle import ransac_oracle as O

    for M, N, s in [(5000, 10000, 1234), (312, 10000, 0), (1 << 19, 1 << 24, 40000), (3125000, 100000000, 777)]:
        a, b = R.estimatescore(M, N, s), O.estimatescore(M, N, s)
        assert (a.min, a.max, a.E) == (b.min, b.max, b.E)


def test_fails_loudly_without_gpu():
    import torch

    import ransac_jl_b200 as R

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(R.RscError) as e:
        R.Context(0)
    assert e.value.code == R._lib.RSC_E_NODEVICE


def test_header_flags_match_the_python_mirror():
    import re

    import ransac_jl_b200 as R

    src = open(os.path.join(ROOT, "include", "rsc.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(RSC_[A-Z_]+)\s+\(?(-?\d+)u?\)?", src)}
    assert defs["RSC_COMPAT_SPHERE_IGNORES_ENABLED"] == R._lib.RSC_COMPAT_SPHERE_IGNORES_ENABLED == 1
    assert defs["RSC_SAMPLER_OCTREE"] == R._lib.RSC_SAMPLER_OCTREE == 2
    assert defs["RSC_REFIT_LSQ"] == R._lib.RSC_REFIT_LSQ == 8
    assert defs["RSC_SCORE_PROGRESSIVE"] == R._lib.RSC_SCORE_PROGRESSIVE == 16
    flags = [defs[k] for k in ("RSC_COMPAT_SPHERE_IGNORES_ENABLED", "RSC_SAMPLER_OCTREE", "RSC_REFIT_LSQ", "RSC_SCORE_PROGRESSIVE")]
    assert len(set(flags)) == 4 and all(f & (f - 1) == 0 for f in flags)  # distinct single bits


def test_estimatescore_f64_mirror_matches_oracle():
    import ransac_jl_b200 as R
    from oracle import ransac_oracle as O

    for M, N, s in [(5000, 10000, 1234), (312, 10000, 0), (1 << 19, 1 << 24, 40000), (3125000, 100000000, 777), (7, 9, 7)]:
        a, b = R.estimatescore_f64(M, N, s), O.estimatescore_f64(M, N, s)
        assert (a.min, a.max, a.E) == (b.min, b.max, b.E)
        assert a.min <= a.E <= a.max


def test_user_defined_shape_brings_its_own_default_parameters():
    # fitting.jl:15 / docs/src/newprimitive.md:12-18
    from dataclasses import dataclass

    import ransac_jl_b200 as R

    @dataclass
    class Torus(R.FittedShape):
        r: float = 1.0

        @staticmethod
        def defaultshapeparameters():
            return {"torus": {"eps": 0.1}}

    p = R.ransacparameters([R.FittedPlane, Torus])
    assert p["torus"] == {"eps": 0.1} and "plane" in p and "sphere" not in p

    class Nothing(R.FittedShape):
        pass

    with pytest.raises(TypeError):
        R.ransacparameters([Nothing])


def test_setfloattype():
    # test/utilitytests.jl:153-189
    import ransac_jl_b200 as R

    nta = {"alpha": 1.0, "somepar": "key1", "intpar": 1}
    ntb = {"alpha": 1, "otherpar": np.float32(0.145), "str": "str"}
    nt = {"alpha": 9, "eps": 0.1, "gamma": np.float32(0.01), "shapea": nta, "shapeb": ntb}
    f32 = R.setfloattype(nt, np.float32)
    assert f32["alpha"] == 9 and isinstance(f32["alpha"], int)
    assert isinstance(f32["eps"], np.float32) and np.isclose(f32["eps"], np.float32(0.1))
    assert isinstance(f32["gamma"], np.float32) and np.isclose(f32["gamma"], np.float32(0.01))
    assert isinstance(f32["shapea"]["alpha"], np.float32) and f32["shapea"]["somepar"] == "key1"
    assert f32["shapea"]["intpar"] == 1 and isinstance(f32["shapea"]["intpar"], int)
    assert f32["shapeb"]["alpha"] == 1 and isinstance(f32["shapeb"]["alpha"], int)
    assert isinstance(f32["shapeb"]["otherpar"], np.float32) and f32["shapeb"]["str"] == "str"
    f64 = R.setfloattype(f32, np.float64)
    rt = float(np.sqrt(np.finfo(np.float32).eps))
    assert f64["alpha"] == 9 and isinstance(f64["eps"], np.float64) and np.isclose(f64["eps"], 0.1, rtol=rt)
    assert isinstance(f64["gamma"], np.float64) and np.isclose(f64["gamma"], 0.01, rtol=rt)
    assert isinstance(f64["shapea"]["alpha"], np.float64) and f64["shapea"]["intpar"] == 1
    assert isinstance(f64["shapeb"]["otherpar"], np.float64) and np.isclose(f64["shapeb"]["otherpar"], 0.145, rtol=rt)
    assert f64["shapeb"]["str"] == "str"


def test_null_handles_are_rejected_without_touching_the_device():
    """every entry point that takes a cloud/context/run returns an error (or a neutral value) for NULL:
    no exception or crash crosses the boundary, and none of this needs a GPU"""
    import ransac_jl_b200 as R

    L = R._lib.lib
    p = R._lib.rsc_params()
    L.rsc_params_default(p)
    cand = R._lib.rsc_cand()
    n64, dbl = C.c_int64(), C.c_double()
    E_ARG = -1
    assert L.rsc_score(None, C.byref(p), None, 0, -1, None, None) == E_ARG
    assert L.rsc_score_culled(None, C.byref(p), None, 0, None, None, None, None) == E_ARG
    assert L.rsc_refit_extract(None, C.byref(p), C.byref(cand), None, C.byref(n64), 0) == E_ARG
    assert L.rsc_refit_lsq(None, C.byref(p), C.byref(cand), 3.0, C.byref(cand), C.byref(n64), C.byref(dbl)) == E_ARG
    assert L.rsc_cloud_build_cells(None, 8) == E_ARG
    assert L.rsc_cloud_set_subset(None, 0, None, 0) == E_ARG
    assert L.rsc_cloud_enable_all(None) == E_ARG
    run = C.c_void_p()
    assert L.rsc_ransac_run(None, C.byref(p), 1, C.byref(run)) == E_ARG and not run.value
    assert L.rsc_run_nshapes(None) == 0 and L.rsc_run_iterations(None) == 0 and L.rsc_run_refined(None) == 0
    assert L.rsc_cloud_size(None) == 0 and L.rsc_cloud_count_enabled(None) == -1
    L.rsc_cloud_destroy(None)
    L.rsc_ctx_destroy(None)
    L.rsc_run_destroy(None)
    assert L.rsc_last_error(None) is not None


def test_update_levelweight_host_function_follows_the_reference_arithmetic():
    """octree.jl:198-205 through the C ABI (a host function: no GPU needed) == the NumPy oracle, bit for bit, and both
    == Julia's evaluation with x = 9//10: Float64(9//10)*sigma/(w*P) + Float64(1//(10 n)) (the rational survives until the sum)"""
    from fractions import Fraction

    import ransac_jl_b200 as R
    from oracle import ransac_oracle as O

    rng = np.random.default_rng(11)
    differs_from_naive = 0
    for n in (1, 3, 7, 8, 9, 11):
        for _ in range(20):
            P = rng.random(n) + 0.05
            P /= P.sum()
            sg = rng.random(n) * rng.integers(0, 2, n) * 1e4
            want = O.updatelevelweight(P.copy(), sg.copy())
            got = P.copy()
            R._lib.lib.rsc_update_levelweight(got.ctypes.data, sg.ctypes.data, n)
            assert np.array_equal(got, want)
            w = 0.0
            for i in range(n):
                w += sg[i] / P[i]
            if w > 0.0:
                julia = np.array([0.9 * sg[i] / (w * P[i]) + float(Fraction(1, 10 * n)) for i in range(n)])
                assert np.array_equal(got, julia)
                naive = np.array([0.9 * sg[i] / (w * P[i]) + (1 - 0.9) / n for i in range(n)])
                differs_from_naive += int(not np.array_equal(naive, julia))
            else:
                assert np.array_equal(got, P)  # no score yet: unchanged
    assert differs_from_naive > 0  # the distinction is real: (1 - 0.9)/n is not the rational's rounding
    cum = np.zeros(8)
    lw = np.full(8, 1 / 8)
    R._lib.lib.rsc_level_cumsum(lw.ctypes.data, 8, cum.ctypes.data)
    assert np.array_equal(cum, O.level_cumsum(lw))
