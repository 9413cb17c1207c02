"""The C oracle's loop (oracle/oracle.c::orc_ransac, restating iterations.jl:35-162) against the NumPy
oracle's loop on the same Philox minimal sets: same shapes in the same order, identical inlier lists
and final isenabled.  (Cone parameters: the NumPy oracle solves the apex with LAPACK, the C oracle by
partial-pivot elimination -- last-ulp differences, rtol 1e-9.)  The C loop is what the GPU tests
compare the device loop with at the c2 / c4 sizes, where the NumPy loop is too slow."""
import numpy as np
import pytest

from oracle import c_oracle
from oracle import ransac_oracle as O

pytestmark = pytest.mark.skipif(not c_oracle.available(), reason="oracle/liboracle.so not built")


def _scene(which):
    from ransac_jl_b200 import scenes

    if which == "c1":
        return scenes.scene_c1(), 2, {}, 4321
    if which == "c1b":
        return scenes.scene_c1(seed=3), 2, {"extract_s": "allcand", "terminate_s": "lengthC", "itermax": 60}, 7
    return (scenes.scene_mixed(81, 20_000, noise_frac=0.002, jitter_deg=1.0, outlier_frac=0.1, counts=(3, 1, 1, 1)), 4,
            {"tau": 200, "minsubsetN": 96, "itermax": 40}, 99)


@pytest.mark.parametrize("which", ["c1", "c1b", "noisy"])
def test_c_loop_equals_numpy_loop(which):
    import ransac_jl_b200 as R
    from tests.helpers import oracle_params

    sc, r, itp, seed = _scene(which)
    op = oracle_params(R.ransacparameters(iteration=itp))
    subs = R.makesubsets(len(sc.vertices), r, np.random.default_rng(1234))
    oc = O.Cloud(sc.vertices, sc.normals, [s.copy() for s in subs])
    tr = O.RansacTrace()
    want = O.ransac(oc, op, True, seed=seed, trace=tr)
    got, en, info = c_oracle.ransac(sc.vertices, sc.normals, subs[0], op, seed)
    assert len(want) == len(got) and len(got) >= 3
    assert info["iterations"] == tr.iterations and info["cands_scored"] == tr.candidates_scored
    assert info["extracted_at"] == tr.extracted_at
    for w, g in zip(want, got):
        assert w.shape.kind == g[0] and (g[0] == 0 or w.shape.outwards == g[1])
        np.testing.assert_allclose(g[2], w.shape.params7(), rtol=1e-9, atol=1e-9)
        np.testing.assert_array_equal(g[3], w.inpoints)
    np.testing.assert_array_equal(en, oc.isenabled)


def test_c_loop_resumes_on_a_partly_disabled_cloud():
    """ransac(pc, params, false): the enabled mask handed in is honoured (iterations.jl:14-21)"""
    import ransac_jl_b200 as R
    from tests.helpers import oracle_params

    sc, r, itp, seed = _scene("c1")
    op = oracle_params(R.ransacparameters(iteration=itp))
    subs = R.makesubsets(len(sc.vertices), r, np.random.default_rng(1234))
    en0 = np.ones(len(sc.vertices), bool)
    en0[:4000] = False  # the plane of c1
    oc = O.Cloud(sc.vertices, sc.normals, [s.copy() for s in subs], en0.copy())
    want = O.ransac(oc, op, False, seed=seed)
    got, en, _ = c_oracle.ransac(sc.vertices, sc.normals, subs[0], op, seed, enabled=en0)
    assert [w.shape.kind for w in want] == [g[0] for g in got]
    for w, g in zip(want, got):
        np.testing.assert_array_equal(g[3], w.inpoints)
        assert g[3].min() >= 4000
    np.testing.assert_array_equal(en, oc.isenabled)


def test_c_estimatescore_equals_numpy_and_the_reference_test():
    """confidenceintervals.jl:53-74 incl. the Int64 wrap (Q9); E is what the loop compares"""
    for m, n, s in [(5000, 10_000, 0), (5000, 10_000, 4999), (312_500, 10_000_000, 25_000), (1_562_500, 100_000_000, 900_000)]:
        ci = O.estimatescore(m, n, s)
        lo, hi, e = c_oracle.estimate_score(m, n, s)
        assert e == pytest.approx(ci.E, rel=1e-15)
        if n <= 10_000:  # no wrap: the whole interval agrees
            assert (lo, hi) == pytest.approx((ci.min, ci.max), rel=1e-13)
