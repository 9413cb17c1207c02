"""End-to-end parity of the device loop (rsc_ransac_run) against the oracle's loop on the same
Philox minimal sets: same extracted shapes in the same order, identical inlier index lists."""
import numpy as np
import pytest

from oracle import ransac_oracle as O
from tests.helpers import oracle_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import ransac_jl_b200 as R

    return R


def _compare_runs(R, extracted, want):
    assert len(extracted) == len(want), ([R.strt(e.shape) for e in extracted], [O.SHAPE_NAMES[w.shape.kind] for w in want])
    for got, w in zip(extracted, want):
        c = got.shape.to_cand()
        assert c.type == w.shape.kind
        if c.type != 0:
            assert bool(c.outwards) == w.shape.outwards
        p = w.shape.params7()
        np.testing.assert_allclose(np.array(c.p[:]), p, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(p).max()))
        np.testing.assert_array_equal(got.inpoints, w.inpoints)


def test_c1_default_parameters_matches_oracle(R):
    """config c1: plane + sphere + cylinder, 10 k points, ransacparameters() defaults, 2 subsets"""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_c1()
    pc = R.RANSACCloud(sc.vertices, sc.normals, 2)
    params = R.ransacparameters()
    extracted, secs = R.ransac(pc, params, True, seed=4321)
    oc = O.Cloud(sc.vertices, sc.normals, [s.copy() for s in pc.subsets])
    tr = O.RansacTrace()
    want = O.ransac(oc, oracle_params(params), True, seed=4321, trace=tr)
    _compare_runs(R, extracted, want)
    kinds = sorted(R.strt(e.shape) for e in extracted)
    assert kinds[:3] == ["cylinder", "plane", "sphere"] or len(kinds) >= 3, kinds
    np.testing.assert_array_equal(pc.isenabled, oc.isenabled)
    print(f"c1: {len(extracted)} shapes {kinds} in {secs} s (oracle: {tr.iterations} iterations, {tr.candidates_scored} candidates)")


def test_noisy_scene_small_matches_oracle(R):
    """noise + outliers, larger minsubsetN, 4 subsets, tau scaled to the cloud"""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(81, 40_000, noise_frac=0.002, jitter_deg=1.0, outlier_frac=0.1, counts=(3, 1, 1, 1))
    pc = R.RANSACCloud(sc.vertices, sc.normals, 4)
    params = R.ransacparameters(iteration={"tau": 400, "minsubsetN": 128, "itermax": 120})
    extracted, secs = R.ransac(pc, params, True, seed=99)
    oc = O.Cloud(sc.vertices, sc.normals, [s.copy() for s in pc.subsets])
    want = O.ransac(oc, oracle_params(params), True, seed=99)
    _compare_runs(R, extracted, want)
    assert len(extracted) >= 3
    np.testing.assert_array_equal(pc.isenabled, oc.isenabled)


def test_resume_without_reset(R):
    """ransac(pc, params, false) continues on the remaining points (iterations.jl:14-21)"""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_c1(seed=3)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 2)
    params = R.ransacparameters(iteration={"itermax": 40})
    first, _ = R.ransac(pc, params, True, seed=7)
    left = pc.count_enabled()
    second, _ = R.ransac(pc, params, False, seed=8)
    assert pc.count_enabled() <= left
    taken = np.concatenate([e.inpoints for e in first + second]) if first or second else np.zeros(0, int)
    assert len(np.unique(taken)) == len(taken)  # no point is extracted twice
    assert pc.count_enabled() == pc.size - len(taken)


@pytest.mark.parametrize("opts", [{}, {"progressive": True}, {"lsq": True}], ids=["plain", "progressive", "lsq"])
def test_sharded_callback_single_rank(R, opts):
    """the sharded code path (range + all-reduce callback) on one rank equals the plain run, also with
    the extensions switched on (progressive scoring slices every subset copy, the least-squares refit
    all-reduces its selection mask)"""
    import ctypes as C

    from ransac_jl_b200 import scenes
    from ransac_jl_b200._lib import ALLREDUCE_FN, lib

    sc = scenes.scene_c1(seed=11)
    params = R.ransacparameters(iteration={"itermax": 120})
    pc = R.RANSACCloud(sc.vertices, sc.normals, 2)
    plain, _ = R.ransac(pc, params, True, seed=5, **opts)
    refined = getattr(pc, "last_refined", 0)
    calls = []
    cb = ALLREDUCE_FN(lambda user, ptr, count, stream: calls.append(count) or 0)
    pc.ctx.check(lib.rsc_ctx_set_allreduce(pc.ctx.h, C.cast(cb, C.c_void_p), None))
    try:
        pc.ctx.check(lib.rsc_cloud_set_range(pc.handle, 0, pc.size))
        sharded, _ = R.ransac(pc, params, True, seed=5, **opts)
    finally:
        lib.rsc_ctx_set_allreduce(pc.ctx.h, None, None)
    assert len(calls) > 0
    assert pc.last_refined == refined and (refined > 0 or "progressive" not in opts)
    assert len(plain) == len(sharded)
    for a, b in zip(plain, sharded):
        assert list(a.shape.to_cand().p) == list(b.shape.to_cand().p)
        np.testing.assert_array_equal(a.inpoints, b.inpoints)


def test_large_cloud_matches_oracle(R):
    """300 k points (the sampler's rank/select index spans more than one scan pass), 8 subsets,
    256 minimal sets per iteration: shapes, inlier lists and the final isenabled equal the oracle's"""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(83, 300_000, noise_frac=0.002, jitter_deg=1.0, outlier_frac=0.1, counts=(3, 1, 1, 1))
    pc = R.RANSACCloud(sc.vertices, sc.normals, 8)
    params = R.ransacparameters(iteration={"tau": 3000, "minsubsetN": 256, "itermax": 40})
    extracted, secs = R.ransac(pc, params, True, seed=99)
    oc = O.Cloud(sc.vertices, sc.normals, [s.copy() for s in pc.subsets])
    want = O.ransac(oc, oracle_params(params), True, seed=99)
    _compare_runs(R, extracted, want)
    assert len(extracted) >= 3
    np.testing.assert_array_equal(pc.isenabled, oc.isenabled)


def test_tiny_guard_queue_and_wide_cones_match_oracle(R):
    """a plane-dominated scene (three-point cone fits on planar patches give near-180-degree cones,
    the kConeWide column type) run on a context whose guard-band queue starts with 32 entries: the
    loop must notice every overflow, grow the queue and repeat, and still equal the oracle"""
    import os

    from ransac_jl_b200 import scenes

    os.environ["RSC_WL_CAP"] = "32"
    try:
        ctx = R.Context(0)
    finally:
        del os.environ["RSC_WL_CAP"]
    sc = scenes.scene_mixed(85, 60_000, noise_frac=0.003, jitter_deg=2.0, outlier_frac=0.2, counts=(4, 1, 0, 0))
    pc = R.RANSACCloud(sc.vertices, sc.normals, 4, ctx=ctx)
    params = R.ransacparameters(iteration={"tau": 600, "minsubsetN": 256, "itermax": 60})
    extracted, secs = R.ransac(pc, params, True, seed=7)
    oc = O.Cloud(sc.vertices, sc.normals, [s.copy() for s in pc.subsets])
    tr = O.RansacTrace()
    want = O.ransac(oc, oracle_params(params), True, seed=7, trace=tr)
    _compare_runs(R, extracted, want)
    assert len(extracted) >= 3
    np.testing.assert_array_equal(pc.isenabled, oc.isenabled)


def test_undersized_sync_free_launch_falls_back_and_matches_oracle(R, monkeypatch):
    """the sync-free batch path sizes its launch by a prediction of the number of new candidates; with the
    prediction forced to 128 (RSC_SMALL_CAP) most batches of this scene hold more, decide_kernel reports it and
    the batch is scored again the classic way -- results must still equal the oracle's loop"""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(87, 60_000, noise_frac=0.002, jitter_deg=1.0, outlier_frac=0.05, counts=(3, 1, 1, 1))
    pc = R.RANSACCloud(sc.vertices, sc.normals, 4)
    params = R.ransacparameters(iteration={"tau": 600, "minsubsetN": 2048, "itermax": 40})
    monkeypatch.setenv("RSC_SMALL_CAP", "128")
    extracted, _ = R.ransac(pc, params, True, seed=3)
    monkeypatch.delenv("RSC_SMALL_CAP")
    plain, _ = R.ransac(pc, params, True, seed=3)
    assert [len(e.inpoints) for e in extracted] == [len(e.inpoints) for e in plain]
    from oracle import c_oracle  # the NumPy loop needs minutes for this many candidates

    want, en, _ = c_oracle.ransac(sc.vertices, sc.normals, pc.subsets[0], oracle_params(params), 3)
    assert len(want) == len(extracted) >= 3
    for got, w in zip(extracted, want):
        assert got.shape.to_cand().type == w[0]
        np.testing.assert_array_equal(got.inpoints, w[3])
    np.testing.assert_array_equal(pc.isenabled, en)


@pytest.mark.parametrize("extract_s,terminate_s", [("allcand", "lengthC"), ("lengthC", "allcand"), ("nofminset", "allcand")])
def test_counter_policies_of_the_decision_kernel_match_oracle(R, extract_s, terminate_s):
    """chooseS (utilities.jl:297-300): the extraction and termination tests may read any of the three counters
    (candidates in the store, candidates ever scored, minimal sets drawn); decide_kernel keeps all three per
    batch iteration -- shapes, inlier lists, isenabled and the iteration count equal the C oracle's loop"""
    from oracle import c_oracle
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(89, 50_000, noise_frac=0.002, jitter_deg=1.0, outlier_frac=0.1, counts=(2, 1, 1, 1))
    pc = R.RANSACCloud(sc.vertices, sc.normals, 4)
    params = R.ransacparameters(iteration={"tau": 500, "minsubsetN": 256, "itermax": 80, "extract_s": extract_s,
                                           "terminate_s": terminate_s})
    extracted, _ = R.ransac(pc, params, True, seed=21)
    want, en, info = c_oracle.ransac(sc.vertices, sc.normals, pc.subsets[0], oracle_params(params), 21)
    assert len(extracted) == len(want)
    for got, w in zip(extracted, want):
        assert got.shape.to_cand().type == w[0]
        np.testing.assert_array_equal(got.inpoints, w[3])
    np.testing.assert_array_equal(pc.isenabled, en)
    assert pc.last_run_iterations == info["iterations"]


def test_drawn_four_points_and_resume_match_oracle(R):
    """drawN = 4 (the fourth point only validates, like the reference's 4-point known-answer sets) through the
    generic fit kernel, then ransac(pc, params, false) on what is left -- both against the C oracle's loop"""
    from oracle import c_oracle
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(90, 40_000, noise_frac=0.001, jitter_deg=0.5, outlier_frac=0.1, counts=(2, 1, 1, 0))
    pc = R.RANSACCloud(sc.vertices, sc.normals, 4)
    params = R.ransacparameters(iteration={"drawN": 4, "tau": 400, "minsubsetN": 512, "itermax": 30})
    op = oracle_params(params)
    first, _ = R.ransac(pc, params, True, seed=31)
    want1, en1, _ = c_oracle.ransac(sc.vertices, sc.normals, pc.subsets[0], op, 31)
    assert len(first) == len(want1) >= 2
    for got, w in zip(first, want1):
        np.testing.assert_array_equal(got.inpoints, w[3])
    np.testing.assert_array_equal(pc.isenabled, en1)
    second, _ = R.ransac(pc, params, False, seed=32)  # continues on the remaining points
    want2, en2, _ = c_oracle.ransac(sc.vertices, sc.normals, pc.subsets[0], op, 32, enabled=en1)
    assert len(second) == len(want2)
    for got, w in zip(second, want2):
        np.testing.assert_array_equal(got.inpoints, w[3])
    np.testing.assert_array_equal(pc.isenabled, en2)
