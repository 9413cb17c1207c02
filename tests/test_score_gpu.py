"""GPU parity of K2 (score) and K4 (refit/extract) against the float64 oracle, through the C ABI.

Bar (BASELINE.json north_star): inlier masks and counts bit-exact, except points within 1e-6
relative of a threshold (none are expected: guard-band pairs are re-evaluated in FP64)."""
import math

import numpy as np
import pytest

from oracle import ransac_oracle as O
from tests.helpers import oracle_mask, oracle_params, to_oracle_shape

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import ransac_jl_b200 as R

    return R


def _check_masks(R, pc, cands, subset_id, params, enabled=None):
    op = oracle_params(params)
    counts, masks = R.score_counts(pc, cands, subset_id, params, want_masks=True)
    counts2, _ = R.score_counts(pc, cands, subset_id, params, want_masks=False)
    np.testing.assert_array_equal(counts, counts2)
    sub = pc.subsets[subset_id] if subset_id >= 0 else np.arange(pc.size)
    pts, nrm = pc.vertices[sub].astype(np.float64), pc.normals[sub].astype(np.float64)
    en = None if enabled is None else enabled[sub]
    nbad = 0
    for i, sh in enumerate(cands):
        honour = sh.kind != R._lib.RSC_SPHERE  # Q4
        want = oracle_mask(sh, pts, nrm, op, en, honour)
        got = R.unpack_mask(masks[i], len(sub))
        assert int(got.sum()) == int(counts[i])
        bad = np.flatnonzero(want != got)
        nbad += len(bad)
        assert len(bad) == 0, f"candidate {i} ({R.strt(sh)}): {len(bad)} mask mismatches at {bad[:8]}"
    return counts


def test_c1_subset_and_whole_cloud(R):
    from ransac_jl_b200 import scenes

    sc = scenes.scene_c1()
    pc = R.RANSACCloud(sc.vertices, sc.normals, 2)
    params = R.ransacparameters()
    cands = [p.shape for p in sc.primitives] + scenes.perturbed_candidates(sc, 8, seed=1)
    counts = _check_masks(R, pc, cands, 0, params)
    assert counts[0] > 1500 and counts[1] > 1000 and counts[2] > 1000  # the true primitives are found
    _check_masks(R, pc, cands, -1, params)
    pc.upload_subset(1)
    _check_masks(R, pc, cands, 1, params)


def test_disabled_points_and_sphere_quirk(R):
    from ransac_jl_b200 import scenes

    sc = scenes.scene_c1(seed=5)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 4)
    rng = np.random.default_rng(0)
    en = rng.random(pc.size) > 0.3
    pc.isenabled = en
    np.testing.assert_array_equal(pc.isenabled, en)
    assert pc.count_enabled() == int(en.sum())
    params = R.ransacparameters()
    cands = [p.shape for p in sc.primitives] + scenes.perturbed_candidates(sc, 4, seed=2)
    _check_masks(R, pc, cands, 0, params, enabled=en)
    _check_masks(R, pc, cands, -1, params, enabled=en)
    # scorecandidate mirrors the reference's return: (ConfidenceInterval, inpoints)
    oc = O.Cloud(sc.vertices, sc.normals, [s.copy() for s in pc.subsets], en.copy())
    for sh in cands[:6]:
        ci, ip = R.scorecandidate(pc, sh, 0, params)
        oci, oip = O.scorecandidate(oc, to_oracle_shape(sh), 0, oracle_params(params))
        np.testing.assert_array_equal(ip, oip)
        assert (ci.min, ci.max, ci.E) == (oci.min, oci.max, oci.E)


def test_mixed_scene_noise_outliers(R):
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(21, 150_000)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 8)
    params = R.ransacparameters()
    cands = [p.shape for p in sc.primitives] + scenes.perturbed_candidates(sc, 12, seed=3)
    _check_masks(R, pc, cands, 0, params)
    counts = _check_masks(R, pc, cands, -1, params)
    assert counts.max() > 3000
    st = pc.ctx.stats()
    print("exact_pairs/evals:", st.exact_pairs, st.evals)


@pytest.mark.parametrize("n", [1, 5, 31, 32, 33, 511, 512, 513, 1025])
def test_ragged_sizes(R, n):
    from ransac_jl_b200 import scenes

    sc = scenes.scene_c1(seed=9)
    pc = R.RANSACCloud(sc.vertices[:n], sc.normals[:n], 1)
    params = R.ransacparameters()
    cands = [p.shape for p in sc.primitives]
    _check_masks(R, pc, cands, 0, params)
    _check_masks(R, pc, cands, -1, params)


def test_degenerate_candidates_and_points(R):
    """NaN parameters match nothing; a point on a sphere centre / cylinder axis / cone apex is
    incompatible (Q17); zero candidates is a no-op."""
    P = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 2], [0, 3, 0], [1, 1, 1]], np.float32)
    N = np.array([[0, 0, 1], [1, 0, 0], [0, 0, 1], [0, 1, 0], [0.57735, 0.57735, 0.57735]], np.float32)
    pc = R.RANSACCloud(P, N, 1)
    params = R.ransacparameters()
    cands = [
        R.FittedSphere([0, 0, 0], 1.0, True),
        R.FittedSphere([0, 0, 0], 0.0, True),
        R.FittedCylinder([0, 0, 1.0], [0, 0, 0], 1.0, True),
        R.FittedCone([0, 0, 0], [0, 0, 1.0], math.radians(60), True),
        R.FittedPlane([0, 0, float("nan")], [0, 0, 1.0]),
        R.FittedSphere([float("inf"), 0, 0], 1.0, True),
        R.FittedPlane([0, 0, 0], [0, 0, 0.0]),
    ]
    _check_masks(R, pc, cands, -1, params)
    c, m = R.score_counts(pc, [], -1, params)
    assert len(c) == 0


def test_thresholds_other_than_default(R):
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(31, 40_000, noise_frac=0.003)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 3)
    params = R.ransacparameters(
        plane={"eps": 0.05, "alpha": math.radians(1)},
        sphere={"eps": 1.0, "alpha": math.radians(20)},
        cylinder={"eps": 0.5, "alpha": math.radians(10)},
        cone={"eps": 0.2, "alpha": math.radians(3)},
    )
    cands = [p.shape for p in sc.primitives] + scenes.perturbed_candidates(sc, 6, seed=4)
    _check_masks(R, pc, cands, 0, params)


def test_refit_extract_matches_oracle(R):
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(41, 100_000)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 4)
    rng = np.random.default_rng(1)
    en = rng.random(pc.size) > 0.2
    pc.isenabled = en
    params = R.ransacparameters()
    op = oracle_params(params)
    oc = O.Cloud(sc.vertices, sc.normals, [s.copy() for s in pc.subsets], en.copy())
    for sh in [p.shape for p in sc.primitives]:
        ex = R.refit(sh, pc, params)
        want = O.refit(to_oracle_shape(sh), oc, op)
        np.testing.assert_array_equal(ex.inpoints, want)
    # extraction disables exactly those points, on the cloud and on the gathered subset copy
    sh = sc.primitives[0].shape
    ex = R.refit(sh, pc, params, disable=True)
    en2 = en.copy()
    en2[ex.inpoints] = False
    np.testing.assert_array_equal(pc.isenabled, en2)
    assert pc.count_enabled() == int(en2.sum())
    counts, masks = R.score_counts(pc, [sh], 0, params, want_masks=True)
    assert counts[0] == 0  # its subset inliers are all disabled now
    # idempotence: a second extraction finds nothing
    assert len(R.refit(sh, pc, params).inpoints) == 0
    pc.enable_all()
    assert pc.count_enabled() == pc.size


def test_large_linearity_and_sample_vs_oracle(R):
    """4096 candidates x 1M points (K=4 path): counts over the whole cloud equal the sum over the
    subsets of a partition; a sample of candidates is checked against the oracle."""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(51, 1 << 20)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 2)
    pc.upload_subset(1)
    params = R.ransacparameters()
    cands = scenes.perturbed_candidates(sc, 1024, seed=6)
    whole, _ = R.score_counts(pc, cands, -1, params)
    s0, _ = R.score_counts(pc, cands, 0, params)
    s1, _ = R.score_counts(pc, cands, 1, params)
    np.testing.assert_array_equal(whole, s0 + s1)
    op = oracle_params(params)
    P64, N64 = sc.vertices.astype(np.float64), sc.normals.astype(np.float64)
    for i in list(range(0, 4096, 331)):
        want = int(O.compatibles(to_oracle_shape(cands[i]), P64, N64, op).sum())
        assert whole[i] == want, (i, R.strt(cands[i]), whole[i], want)
    st = pc.ctx.stats()
    print("exact_pairs:", st.exact_pairs, "evals:", st.evals, "kernel ms:", st.last_kernel_ms)


def test_estimatescore_matches_oracle(R):
    for M, N, s in [(5000, 10000, 1234), (312, 10000, 0), (1 << 19, 1 << 24, 40000), (3125000, 100000000, 777)]:
        a, b = R.estimatescore(M, N, s), O.estimatescore(M, N, s)
        assert (a.min, a.max, a.E) == (b.min, b.max, b.E) or (a.E != a.E and b.E != b.E)
