"""Progressive subset scoring: the refinement the reference leaves as
"TODO: refine if best.overlap" (iterations.jl:110; docs/src/ransac.md:137-141; Schnabel 2007 sec. 4.5.1).
CPU: properties of the oracle's refinement.  GPU: the device loop (RSC_SCORE_PROGRESSIVE) and the
package's host loop over the per-call C ABI (rsc_score on subsets 2..r, counts only) both make the
oracle's decisions."""
import numpy as np
import pytest

from oracle import ransac_oracle as O
from tests.helpers import oracle_params


def _scene(n=20_000):
    from ransac_jl_b200 import scenes

    return scenes.scene_mixed(86, n, noise_frac=0.004, jitter_deg=1.5, outlier_frac=0.2, counts=(2, 1, 1, 1))


def test_estimatescore_f64_equals_the_reference_interval_where_int64_does_not_wrap():
    for M, N, s in [(312, 10_000, 0), (5000, 10_000, 3000), (1000, 200_000, 999), (7, 50, 7)]:
        a, b = O.estimatescore(M, N, s), O.estimatescore_f64(M, N, s)
        assert abs(a.min - b.min) <= 1e-9 * max(1, abs(a.min)) and abs(a.max - b.max) <= 1e-9 * max(1, abs(a.max))
    # where it wraps (Q9) E is still the same and the f64 interval is a real interval around it
    a, b = O.estimatescore(500_000, 16_000_000, 123_456), O.estimatescore_f64(500_000, 16_000_000, 123_456)
    assert abs(a.E - b.E) <= 1e-6 * b.E and b.min < b.E < b.max and (b.max - b.min) < 0.02 * b.E


def test_refinement_separates_or_exhausts_and_narrows_intervals():
    sc = _scene()
    rng = np.random.default_rng(5)
    subs = O.make_subsets(len(sc.vertices), 8, rng.permutation(len(sc.vertices)))
    P = O.ransacparameters(O.default_parameters(), iteration={"tau": 300, "minsubsetN": 48, "itermax": 1})
    pc = O.Cloud(sc.vertices.astype(np.float64), sc.normals.astype(np.float64), subs)
    en = np.flatnonzero(pc.isenabled)
    shapes = []
    for i in range(2000):
        ok, _, sd = O.sample_minimal_set(pc, 3, O.SetStream(3, i), en)
        if ok:
            shapes.extend(O.forcefit(pc.vertices[sd], pc.normals[sd], P))
    assert len(shapes) > 10
    M1 = len(subs[0])
    first = [int(O.scorecandidate(pc, s, 0, P)[1].size) for s in shapes]
    scores = [O.estimatescore_f64(M1, pc.size, c) for c in first]
    evaluated = [[1, c, M1] for c in first]
    width0 = [s.max - s.min for s in scores]
    tr = O.RansacTrace()
    O.refine_progressive(pc, P, shapes, scores, evaluated, tr)
    best, overlap = O.findhighestscore(scores)
    group = [i for i in range(len(scores)) if i == best or O.isoverlap(scores[i], scores[best])]
    assert (not overlap) or min(evaluated[i][0] for i in group) == len(subs)
    assert tr.refined == sum(e[0] - 1 for e in evaluated) > 0
    for i, e in enumerate(evaluated):
        assert e[2] == sum(len(subs[j]) for j in range(e[0]))
        # cumulative count = compatible points of the union of the evaluated subsets
        u = np.concatenate(subs[: e[0]])
        cp = O.compatibles(shapes[i], pc.vertices[u], pc.normals[u], P)
        assert int(cp.sum()) == e[1]
        if e[0] == len(subs):  # everything seen: the estimate is the exact count, the interval a point
            assert abs(scores[i].E - e[1]) < 1e-6 and scores[i].max - scores[i].min < 1e-3
        if e[0] > 1 and e[1] > 50:  # relative width shrinks with the sample
            assert (scores[i].max - scores[i].min) / scores[i].E < width0[i] / max(1e-9, O.estimatescore_f64(M1, pc.size, first[i]).E) + 1e-9
    # candidates never looked at again were clearly below the best
    for i, e in enumerate(evaluated):
        if e[0] == 1 and i != best:
            assert scores[i].max < scores[best].min or min(evaluated[j][0] for j in group) == len(subs)


def test_progressive_loop_differs_only_by_better_informed_extractions():
    sc = _scene()
    rng = np.random.default_rng(5)
    subs = O.make_subsets(len(sc.vertices), 8, rng.permutation(len(sc.vertices)))
    P = O.ransacparameters(O.default_parameters(), iteration={"tau": 300, "minsubsetN": 48, "itermax": 40})
    runs = []
    for prog in (False, True):
        pc = O.Cloud(sc.vertices.astype(np.float64), sc.normals.astype(np.float64), subs)
        tr = O.RansacTrace()
        runs.append((O.ransac(pc, P, True, seed=21, trace=tr, progressive=prog), tr))
    (a, ta), (b, tb) = runs
    assert ta.refined == 0 and tb.refined > 0
    assert len(a) >= 3 and len(b) >= 3
    # both find the three big primitives first; the refined run never extracts a smaller first shape
    assert len(b[0].inpoints) >= len(a[0].inpoints)


@pytest.mark.gpu
@pytest.mark.parametrize("loop", ["device", "host"])
def test_progressive_loop_matches_oracle(loop):
    import ransac_jl_b200 as R
    from ransac_jl_b200 import iterations as IT

    sc = _scene(30_000)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 8)
    params = R.ransacparameters(iteration={"tau": 300, "minsubsetN": 48, "itermax": 40})
    if loop == "device":
        extracted, _ = R.ransac(pc, params, True, seed=21, progressive=True)
    else:
        pc.enable_all()
        extracted, _ = IT._ransac_host(pc, params, 21, progressive=True)
    oc = O.Cloud(sc.vertices.astype(np.float64), sc.normals.astype(np.float64), [s.copy() for s in pc.subsets])
    tr = O.RansacTrace()
    want = O.ransac(oc, oracle_params(params), True, seed=21, trace=tr, progressive=True)
    assert tr.refined > 0 and pc.last_refined == tr.refined
    assert len(extracted) == len(want) >= 3
    for got, w in zip(extracted, want):
        c = got.shape.to_cand()
        assert c.type == w.shape.kind
        p = w.shape.params7()
        np.testing.assert_allclose(np.array(c.p[:7]), p, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(p).max()))
        np.testing.assert_array_equal(got.inpoints, w.inpoints)
    np.testing.assert_array_equal(pc.isenabled, oc.isenabled)
