"""Pins the NumPy oracle to the reference's own known-answer tests.

Each case restates a testset of /root/reference/test (file:line cited per test) with the
same inputs and the same expected outcome.
"""
import math

import numpy as np
import pytest

from oracle import ransac_oracle as O

EPSI = 0.1
ALFI = math.radians(10)


def _defrp():
    # test/dummyspheretest.jl:8-10
    return O.ransacparameters(O.ransacparameters(), sphere={"eps": EPSI, "alpha": ALFI})


def _plane_rp(base):
    # test/dummyspheretest.jl:19
    return O.ransacparameters(base, plane={"alpha": math.pi / 2}, common={"collin_threshold": 0.2})


TN1 = [(0, -1, 0.0), (0, 0, -1.0), (1, 0, 0.0), (0, 1, 0.0)]


def test_true_sphere_1():
    # test/dummyspheretest.jl:14-22
    tv = [(0, -1, 0.0), (0, 0, -1.0), (1, 0, 0.0), (0, 1, 0.0)]
    fs = O.fit_sphere(tv, TN1, _defrp())
    fp = O.fit_plane(tv, TN1, _plane_rp(_defrp()))
    assert fs is not None and fs.kind == O.SPHERE
    assert fp is None
    np.testing.assert_allclose(fs.a, [0, 0, 0], atol=1e-15)
    assert fs.s == pytest.approx(1.0, abs=1e-15)
    assert fs.outwards is True


def test_true_sphere_2():
    # test/dummyspheretest.jl:24-35
    tv = [(0, -0.99, 0.0), (0, 0, -1.0), (1.01, 0, 0.0), (0, 1, 0.0)]
    fs1 = O.fit_sphere(tv, TN1, _defrp())
    fs2 = O.fit_sphere(tv, TN1, O.ransacparameters(_defrp(), sphere={"eps": 0.01}))
    fp = O.fit_plane(tv, TN1, _plane_rp(_defrp()))
    assert fs1 is not None and fs1.kind == O.SPHERE
    assert fs2 is None
    assert fp is None


def test_false_sphere_1():
    # test/dummyspheretest.jl:37-49
    tv = [(0, 1, 0.0), (0, 0, -1.0), (1, 0, 0.0), (0, 1, 0.0)]
    fs1 = O.fit_sphere(tv, TN1, _defrp())
    fs2 = O.fit_sphere(tv, TN1, O.ransacparameters(_defrp(), sphere={"eps": 10, "alpha": math.pi / 2}))
    fp = O.fit_plane(tv, TN1, _plane_rp(_defrp()))
    assert fs1 is None and fs2 is None and fp is None


def test_confidence_interval():
    # test/confidenceintervals.jl:1-10
    ci = O.ConfidenceInterval(1.0, 3)
    assert ci.E == 2.0 and ci.min == 1.0 and ci.max == 3.0
    with pytest.raises(ValueError):
        O.ConfidenceInterval(3, 1.0)


def test_notsoconfident():
    # test/confidenceintervals.jl:12-25
    nc1 = O.notsoconfident(153.9, 9.7)
    nc2 = O.notsoconfident(9.7, 153.9)
    assert nc1.min == 9.7 and nc1.max == 153.9 and nc1.E == 81.8
    assert (nc1.min, nc1.max, nc1.E) == (nc2.min, nc2.max, nc2.E)


def test_default_parameters():
    # test/utilitytests.jl:41-82
    p = O.default_parameters()
    it = p["iteration"]
    assert (it["drawN"], it["minsubsetN"], it["prob_det"], it["tau"], it["itermax"]) == (3, 15, 0.9, 900, 1000)
    assert it["extract_s"] == "nofminset" and it["terminate_s"] == "nofminset"
    assert it["shape_types"] == [O.PLANE, O.CONE, O.CYLINDER, O.SPHERE]  # src/RANSAC.jl:94
    assert p["common"] == {"collin_threshold": 0.2, "parallelthrdeg": 1.0}
    assert p["plane"] == {"eps": 0.3, "alpha": math.radians(5)}
    assert p["sphere"] == {"eps": 0.3, "alpha": math.radians(5), "sphere_par": 0.02}
    assert p["cylinder"] == {"eps": 0.3, "alpha": math.radians(5)}
    assert p["cone"] == {"eps": 0.3, "alpha": math.radians(5), "minconeopang": math.radians(2)}


def test_ransacparameters_merge():
    # test/utilitytests.jl:84-114
    p = O.ransacparameters(O.default_parameters([O.SPHERE, O.CYLINDER]), sphere={"eps": 0.01}, cylinder={"alpha": 0.02})
    assert p["sphere"] == {"eps": 0.01, "alpha": math.radians(5), "sphere_par": 0.02}
    assert p["cylinder"] == {"eps": 0.3, "alpha": 0.02}


def test_estimatescore_closed_form():
    # E = -1 + (N+2)(sigma+1)/(M+2)  (SURVEY 8a22); interval ordered; small sizes, no overflow
    for M, N, s in [(5000, 10000, 1234), (312, 10000, 0), (312, 10000, 312), (100, 100, 50)]:
        ci = O.estimatescore(M, N, s)
        assert ci.min <= ci.max
        assert ci.E == pytest.approx(-1 + (N + 2) * (s + 1) / (M + 2), rel=1e-12)


def test_estimatescore_int64_overflow_keeps_E():
    # Q9: the Int64 product wraps for N >~ 1e6 but E is unaffected
    M, N, s = 1 << 19, 1 << 24, 40000
    ci = O.estimatescore(M, N, s)
    assert ci.E == pytest.approx(-1 + (N + 2) * (s + 1) / (M + 2), rel=1e-12)


def test_octree_depth_on_the_reference_grid():
    # test/octree.jl:116-139: 6^3 grid with spacing 1/3 -> every leaf has depth 3, the tree has depth 3,
    # getnthcell(leaf, 2) / (leaf, 1) are the parent / the root.  Restated on the flattened octree.
    ps = np.array([[i, j, k] for i in range(6) for j in range(6) for k in range(6)], float) / 3
    for nlevels in (3, 5, 8):
        oc = O.MortonOctree(ps, nlevels)
        assert (oc.leafdepth == 3).all()  # findleaf(pc.octree, ps[117]).data.depth == 3; octreedepth == 3
        a3, b3 = oc.cell_range(116, 3)  # ps[117] is 1-based
        a2, b2 = oc.cell_range(116, 2)
        a1, b1 = oc.cell_range(116, 1)
        assert (a1, b1) == (0, 216) and b2 - a2 == 27 and 1 <= b3 - a3 <= 8
        assert a1 <= a2 <= a3 and b3 <= b2 <= b1  # leaf within parent within root
        assert set(oc.perm[a3:b3]) <= set(oc.perm[a2:b2])


def test_cross_product_tensor_and_rodrigues():
    """test/utilitytests.jl:115-133 ("cross prod"): pluscrossprod!(A, value, v) == A + value * crossprodtensor(v),
    exactly; plus the defining properties of the rotation the cone projection builds from them"""
    import math

    rng = np.random.default_rng(0)
    A = rng.random((3, 3))
    val = math.radians(15)
    vec = rng.random(3)
    vec /= np.linalg.norm(vec)
    assert np.array_equal(A + math.sin(val) * O.crossprodtensor(vec), O.pluscrossprod(A.copy(), math.sin(val), vec))
    assert np.array_equal(A + O.crossprodtensor(vec), O.pluscrossprod(A.copy(), 1, vec))
    assert np.array_equal(A, O.pluscrossprod(A.copy(), 0, vec))
    # rodrigues (utilities.jl:6-11): a rotation about vec by val
    R = O.rodrigues(vec[None, :], math.cos(val), math.sin(val))[0]
    np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-15)
    np.testing.assert_allclose(R @ vec, vec, atol=1e-15)
    assert abs(np.trace(R) - (1 + 2 * math.cos(val))) < 1e-15 and abs(np.linalg.det(R) - 1) < 1e-15
    perp = np.cross(vec, [1.0, 0, 0])
    perp /= np.linalg.norm(perp)
    assert np.dot(np.cross(perp, R @ perp), vec) > 0  # right-handed about vec


def test_push2candidatesandlevels():
    """test/utilitytests.jl:135-151"""
    fp = O.Shape(O.PLANE, np.array([0.5, 0.5, 0.5]), np.array([0.0, 0, 1]))
    candidates, levels = [], []
    O.push2candidatesandlevels(candidates, fp, levels, 3)
    assert len(candidates) == 1 and len(levels) == 1
    O.push2candidatesandlevels(candidates, [fp, fp], levels, 0)
    assert len(candidates) == 3 and levels == [3, 0, 0]


def test_iswithinrectangle():
    """test/octree.jl:10-114 on the unit box: corner points, edge / face midpoints, inside and outside points"""
    lo, hi = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)
    inside = [(1, 1, 1), (1, .5, 1), (.5, 1, 1), (1, 1, .5), (.5, .5, 1), (.5, 1, .5), (1, .5, .5),
              (.5, .5, .1), (.5, .1, .5), (.1, .5, .5), (.5, .5, .9), (.5, .9, .5), (.9, .5, .5)]
    outside = [(0, 0, 0), (1, 0, 0), (0, 1, 0), (0, 0, 1), (1, 1, 0), (1, 0, 1), (0, 1, 1),
               (.5, 0, 0), (.5, 0, 1), (0, .5, 0), (0, .5, 1), (1, .5, 0), (.5, 1, 0), (1, 0, .5), (0, 0, .5), (0, 1, .5),
               (.5, .5, 0), (.5, 0, .5), (0, .5, .5),
               (.5, .5, -.1), (.5, -.1, .5), (-.1, .5, .5), (.5, .5, 1.1), (.5, 1.1, .5), (1.1, .5, .5)]
    for p in inside:
        assert O.iswithinrectangle(lo, hi, p) is True, p
    for p in outside:
        assert O.iswithinrectangle(lo, hi, p) is False, p


def test_findAABB():
    """test/utilitytests.jl:5-27: 30 random points in the unit cube/square plus the two extreme corners"""
    rng = np.random.default_rng(1)
    for dim in (3, 2):
        pts = list(rng.random((30, dim))) + [np.full(dim, -1.0), np.full(dim, 2.0)]
        mn, mx = O.findAABB(pts)
        assert np.array_equal(mn, np.full(dim, -1.0)) and np.array_equal(mx, np.full(dim, 2.0))


def test_smallestdistance():
    """test/utilitytests.jl:29-39 (off the hot path; pinned for completeness)"""
    assert np.isclose(O.smallestdistance([(0.0, 0), (1.0, 1), (2.2, 2)]), np.sqrt(2))
    assert np.isclose(O.smallestdistance([(0.0, 0, 0), (1.0, 1, 1), (2.2, 2, 2)]), np.sqrt(3))
