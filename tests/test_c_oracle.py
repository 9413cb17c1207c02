"""The C restatement (oracle/oracle.c) against the NumPy restatement (oracle/ransac_oracle.py):
identical inlier masks bit for bit, identical fit accept/reject and parameters, and the reference's
known answers (test/dummyspheretest.jl) through the C code as well."""
import math

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import ransac_oracle as O

pytestmark = pytest.mark.skipif(not CO.available(), reason="oracle/liboracle.so not built (run __graft_entry__.build())")


def _scene(seed, n):
    import ransac_jl_b200.scenes as S

    return S.scene_mixed(seed, n)


def test_masks_identical_to_numpy_oracle():
    import ransac_jl_b200.scenes as S
    from tests.helpers import to_oracle_shape

    sc = _scene(5, 30_000)
    P, N = sc.vertices.astype(np.float64), sc.normals.astype(np.float64)
    cands = [p.shape for p in sc.primitives] + S.perturbed_candidates(sc, 10, seed=2)
    op = O.default_parameters()
    en = np.random.default_rng(0).random(len(P)) > 0.4
    counts, _, masks = CO.score_counts(cands, P, N, op, enabled=en, want_masks=True)
    for i, sh in enumerate(cands):
        want = O.compatibles(to_oracle_shape(sh), P, N, op)
        if sh.kind != 1:
            want = want & en
        assert np.array_equal(masks[i], want), (i, sh)
        assert counts[i] == want.sum()


def test_fits_match_numpy_oracle():
    rng = np.random.default_rng(7)
    sc = _scene(6, 20_000)
    P, N = sc.vertices.astype(np.float64), sc.normals.astype(np.float64)
    S_ = 3000
    idx = np.empty((S_, 3), np.int64)
    # half the sets from within one primitive (so that fits succeed), half random
    lab = sc.labels
    for s in range(S_):
        if s % 2 == 0:
            l = rng.integers(0, lab.max() + 1)
            pool = np.flatnonzero(lab == l)
            idx[s] = rng.choice(pool, 3, replace=False)
        else:
            idx[s] = rng.choice(len(P), 3, replace=False)
    op = O.default_parameters()
    got, got_set = CO.fit_points(P[idx], N[idx], op)
    want = []
    for s in range(S_):
        for sh in O.forcefit(P[idx[s]], N[idx[s]], op):
            want.append((s, sh))
    assert len(got) == len(want), (len(got), len(want))
    ntype = [0, 0, 0, 0]
    for (t, outw, p), s, (ws, wsh) in zip(got, got_set, want):
        assert s == ws and t == wsh.kind and outw == wsh.outwards
        np.testing.assert_allclose(p, wsh.params7(), rtol=1e-6, atol=1e-6)  # 1e-5 is the bar; ill-conditioned cone apex solves differ ~1e-8 between LAPACK and plain elimination
        ntype[t] += 1
    assert min(ntype) > 20, ntype  # every shape type was exercised


def test_known_answers_through_c():
    # test/dummyspheretest.jl:14-48
    tn = np.array([(0, -1, 0.0), (0, 0, -1.0), (1, 0, 0.0), (0, 1, 0.0)])
    base = O.ransacparameters(O.default_parameters([O.SPHERE]), sphere={"eps": 0.1, "alpha": math.radians(10)})
    tv1 = np.array([(0, -1, 0.0), (0, 0, -1.0), (1, 0, 0.0), (0, 1, 0.0)])
    tv2 = np.array([(0, -0.99, 0.0), (0, 0, -1.0), (1.01, 0, 0.0), (0, 1, 0.0)])
    tv3 = np.array([(0, 1, 0.0), (0, 0, -1.0), (1, 0, 0.0), (0, 1, 0.0)])
    r, _ = CO.fit_points(tv1[None], tn[None], base)
    assert len(r) == 1 and r[0][0] == 1 and r[0][1] is True
    np.testing.assert_allclose(r[0][2][:4], [0, 0, 0, 1], atol=1e-15)
    assert len(CO.fit_points(tv2[None], tn[None], base)[0]) == 1
    assert len(CO.fit_points(tv2[None], tn[None], O.ransacparameters(base, sphere={"eps": 0.01}))[0]) == 0
    assert len(CO.fit_points(tv3[None], tn[None], base)[0]) == 0
    assert len(CO.fit_points(tv3[None], tn[None], O.ransacparameters(base, sphere={"eps": 10, "alpha": math.pi / 2}))[0]) == 0
    pl = O.ransacparameters(O.default_parameters([O.PLANE]), plane={"alpha": math.pi / 2})
    for tv in (tv1, tv2, tv3):
        assert len(CO.fit_points(tv[None], tn[None], pl)[0]) == 0


def test_random_and_degenerate_candidates_agree_between_the_two_oracles():
    """adversarial inputs: random shapes of every type (wide/flat cones, non-unit axes and normals,
    huge and tiny radii), points on axes / at centres / with zero normals, NaN parameters -- the two
    independent restatements (NumPy, C) must still give the same masks"""
    from tests.helpers import adversarial_case

    shapes, P, N = adversarial_case()
    op = O.default_parameters()
    counts, _, masks = CO.score_counts(shapes, P, N, op, want_masks=True)
    nonzero = 0
    for i, sh in enumerate(shapes):
        with np.errstate(all="ignore"):
            want = O.compatibles(sh, P, N, op)
        assert np.array_equal(masks[i], want), (i, sh.kind, int(want.sum()), int(masks[i].sum()))
        nonzero += int(want.any())
    assert nonzero > 20


def test_degenerate_minimal_sets_agree_between_the_two_oracles():
    """collinear / coincident points, parallel, anti-parallel and zero normals, near-singular cone normal
    triples (rank() decisions at their tolerance), scales 1e-3..1e3: same accept/reject, same parameters"""
    from tests.helpers import degenerate_sets

    P, N = degenerate_sets()
    op = O.default_parameters()
    got, gset = CO.fit_points(P, N, op)
    want = []
    with np.errstate(all="ignore"):
        for s in range(len(P)):
            want += [(s, sh) for sh in O.forcefit(P[s], N[s], op)]
    assert [(int(s), t) for (t, _, _), s in zip(got, gset)] == [(s, sh.kind) for s, sh in want]
    assert len(want) > 40
    for (t, outw, p), (_, sh) in zip(got, want):
        assert t == 0 or bool(outw) == bool(sh.outwards)
        np.testing.assert_allclose(np.asarray(p)[:7], sh.params7(), rtol=1e-9, atol=1e-9 * max(1.0, np.abs(sh.params7()).max()))
