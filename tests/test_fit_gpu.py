"""GPU parity of K1 (batched fits, sampler) against the CPU restatements, through the C ABI.

Bar: identical accept/reject and candidate order; fitted parameters within 1e-5 relative
(BASELINE.json north_star)."""
import math

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import ransac_oracle as O
from tests.helpers import oracle_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import ransac_jl_b200 as R

    return R


def _sets(sc, S, seed):
    """half the sets inside one primitive (fits succeed), half random"""
    rng = np.random.default_rng(seed)
    idx = np.empty((S, 3), np.int64)
    lab = sc.labels
    pools = [np.flatnonzero(lab == l) for l in range(lab.max() + 1)]
    for s in range(S):
        if s % 2 == 0:
            idx[s] = rng.choice(pools[rng.integers(len(pools))], 3, replace=False)
        else:
            idx[s] = rng.choice(len(lab), 3, replace=False)
    return idx


def _compare(R, shapes, sets, want, want_set):
    assert len(shapes) == len(want), (len(shapes), len(want))
    np.testing.assert_array_equal(sets, want_set)
    ntype = [0, 0, 0, 0]
    for sh, (t, outw, p) in zip(shapes, want):
        c = sh.to_cand()
        assert c.type == t and bool(c.outwards) == outw
        got = np.array(c.p[:])
        scale = max(1.0, float(np.abs(p).max()))
        assert np.abs(got - p).max() <= 1e-5 * scale, (t, got, p)
        ntype[t] += 1
    return ntype


def test_known_answers_dummyspheretest(R):
    # test/dummyspheretest.jl:14-48 through fit(::Type{S}, p, n, pc, params)
    pc = R.RANSACCloud(np.random.default_rng(0).random((10, 3)), np.random.default_rng(1).random((10, 3)), 1)
    defrp = R.ransacparameters(R.ransacparameters(), sphere={"eps": 0.1, "alpha": math.radians(10)})
    plane_rp = R.ransacparameters(defrp, plane={"alpha": math.pi / 2}, common={"collin_threshold": 0.2})
    tn = [(0, -1, 0.0), (0, 0, -1.0), (1, 0, 0.0), (0, 1, 0.0)]
    tv1 = [(0, -1, 0.0), (0, 0, -1.0), (1, 0, 0.0), (0, 1, 0.0)]
    tv2 = [(0, -0.99, 0.0), (0, 0, -1.0), (1.01, 0, 0.0), (0, 1, 0.0)]
    tv3 = [(0, 1, 0.0), (0, 0, -1.0), (1, 0, 0.0), (0, 1, 0.0)]
    fs = R.fit(R.FittedSphere, tv1, tn, pc, defrp)
    assert isinstance(fs, R.FittedSphere) and fs.outwards
    np.testing.assert_allclose(fs.center, [0, 0, 0], atol=1e-12)
    assert abs(fs.radius - 1.0) < 1e-12
    assert R.fit(R.FittedPlane, tv1, tn, pc, plane_rp) is None
    assert isinstance(R.fit(R.FittedSphere, tv2, tn, pc, defrp), R.FittedSphere)
    assert R.fit(R.FittedSphere, tv2, tn, pc, R.ransacparameters(defrp, sphere={"eps": 0.01})) is None
    assert R.fit(R.FittedPlane, tv2, tn, pc, plane_rp) is None
    assert R.fit(R.FittedSphere, tv3, tn, pc, defrp) is None
    assert R.fit(R.FittedSphere, tv3, tn, pc, R.ransacparameters(defrp, sphere={"eps": 10, "alpha": math.pi / 2})) is None
    assert R.fit(R.FittedPlane, tv3, tn, pc, plane_rp) is None
    with pytest.raises(AssertionError):  # "At least 3 point is needed." (plane.jl:39)
        R.fit_points(pc, np.zeros((1, 3, 3))[:, :2], np.zeros((1, 3, 3))[:, :2], defrp)


def test_fit_batch_matches_oracle(R):
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(61, 60_000)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 4)
    params = R.ransacparameters()
    op = oracle_params(params)
    idx = _sets(sc, 6000, 3)
    P, N = sc.vertices.astype(np.float64), sc.normals.astype(np.float64)
    want, want_set = CO.fit_points(P[idx], N[idx], op)
    shapes, sets = R.fit_batch(pc, idx, params)
    ntype = _compare(R, shapes, sets, want, want_set)
    assert min(ntype) > 30, ntype
    # explicit-coordinate entry gives the same candidates
    shapes2, sets2 = R.fit_points(pc, P[idx], N[idx], params)
    _compare(R, shapes2, sets2, want, want_set)
    # spot-check against the NumPy restatement (LAPACK rank / solve)
    k = 0
    for s in range(0, 400):
        for sh in O.forcefit(P[idx[s]], N[idx[s]], op):
            c = shapes[k].to_cand()
            assert sets[k] == s and c.type == sh.kind and bool(c.outwards) == sh.outwards
            np.testing.assert_allclose(np.array(c.p[:]), sh.params7(), rtol=1e-5, atol=1e-5)
            k += 1


def test_fit_noise_free_degenerate_sets(R):
    """noise-free c1 scene: identical normals on the plane (rank-deficient cone system), exact
    spheres/cylinders; shape_types order other than the default"""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_c1()
    pc = R.RANSACCloud(sc.vertices, sc.normals, 2)
    params = R.ransacparameters([R.FittedSphere, R.FittedPlane, R.FittedCylinder, R.FittedCone])
    op = oracle_params(params)
    idx = _sets(sc, 3000, 5)
    P, N = sc.vertices.astype(np.float64), sc.normals.astype(np.float64)
    want, want_set = CO.fit_points(P[idx], N[idx], op)
    shapes, sets = R.fit_batch(pc, idx, params)
    ntype = _compare(R, shapes, sets, want, want_set)
    assert ntype[0] > 50 and ntype[1] > 50 and ntype[2] > 50, ntype


def test_sampler_matches_oracle_philox(R):
    """rsc_sample_fit draws the same minimal sets as the oracle's Philox sampler (same seed), with
    a third of the points disabled, and fits them like fit_batch."""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(71, 30_000)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 2)
    en = np.random.default_rng(5).random(pc.size) > 0.33
    pc.isenabled = en
    params = R.ransacparameters()
    S, seed, set0 = 2000, 987654321, 12345
    shapes, sets, idx = R.sample_fit(pc, params, seed, set0, S)
    oc = O.Cloud(sc.vertices, sc.normals, [s.copy() for s in pc.subsets], en.copy())
    en_idx = np.flatnonzero(en)
    nfail = 0
    for s in range(S):
        ok, _, want = O.sample_minimal_set(oc, 3, O.SetStream(seed, set0 + s), en_idx)
        if ok:
            np.testing.assert_array_equal(idx[s], want)
            assert en[idx[s]].all()
        else:
            nfail += 1
            assert (idx[s] == -1).all()
    good = np.flatnonzero(idx[:, 0] >= 0)
    shapes2, sets2 = R.fit_batch(pc, idx[good], params)
    assert len(shapes) == len(shapes2)
    np.testing.assert_array_equal(sets, good[sets2])
    for a, b in zip(shapes, shapes2):
        assert list(a.to_cand().p) == list(b.to_cand().p)
    print("failed samples:", nfail, "candidates:", len(shapes))


def test_sampler_too_few_enabled(R):
    from ransac_jl_b200 import scenes

    sc = scenes.scene_c1()
    pc = R.RANSACCloud(sc.vertices[:100], sc.normals[:100], 1)
    en = np.zeros(100, bool)
    en[[3, 50]] = True
    pc.isenabled = en
    shapes, sets, idx = R.sample_fit(pc, R.ransacparameters(), 1, 0, 64)
    assert len(shapes) == 0 and (idx == -1).all()


def test_degenerate_minimal_sets_match_oracle(R):
    """tests.helpers.degenerate_sets through rsc_fit_points: collinear / coincident points, parallel and zero
    normals, near-singular cone normal triples (the rank shortcut must agree with the SVD), three scales"""
    from oracle import c_oracle as CO
    from tests.helpers import degenerate_sets, oracle_params

    P, N = degenerate_sets()
    params = R.ransacparameters()
    pc = R.RANSACCloud(np.zeros((4, 3), np.float32), np.ones((4, 3), np.float32), 1)
    shapes, sets = R.fit_points(pc, P, N, params)
    want, wset = CO.fit_points(P, N, oracle_params(params))
    assert [(int(s), sh.to_cand().type) for s, sh in zip(sets, shapes)] == [(int(s), t) for (t, _, _), s in zip(want, wset)]
    for sh, (t, outw, p) in zip(shapes, want):
        c = sh.to_cand()
        assert t == 0 or bool(c.outwards) == bool(outw)
        p = np.asarray(p)[:7]
        np.testing.assert_allclose(np.array(c.p[:7]), p, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(p).max()))
