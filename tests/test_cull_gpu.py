"""Morton-tile culling (rsc_score_culled): the counts must be the dense path's, which are the oracle's.
Both ways of deciding the in-band pairs are covered: queued for cull_fix_kernel (default) and inline
(RSC_CULL_INLINE=1, read by the library on every call)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import ransac_jl_b200 as R

    return R


@pytest.mark.parametrize("inline", ["0", "1"], ids=["queued", "inline"])
def test_culled_counts_equal_dense_counts_on_a_noisy_scene(R, monkeypatch, inline):
    from ransac_jl_b200 import scenes

    monkeypatch.setenv("RSC_CULL_INLINE", inline)

    sc = scenes.scene_mixed(91, 300_000, noise_frac=0.004, jitter_deg=1.5, outlier_frac=0.2, counts=(3, 2, 2, 2))
    pc = R.RANSACCloud(sc.vertices, sc.normals, 1)
    cands = scenes.perturbed_candidates(sc, 160, seed=7)
    params = R.ransacparameters()
    with pytest.raises(R.RscError):
        R.score_counts_culled(pc, cands, params)  # needs the Morton order
    pc.build_cells(9)
    en = np.ones(pc.size, bool)
    en[np.random.default_rng(3).random(pc.size) < 0.3] = False
    pc.isenabled = en
    dense, _ = R.score_counts(pc, cands, -1, params)
    got, info = R.score_counts_culled(pc, cands, params)
    np.testing.assert_array_equal(got, dense)
    assert info["pairs_total"] == len(cands) * (((pc.size + 511) // 512) * 4)  # 128-point tiles of the padded cloud
    assert 0 < info["pairs_survived"] < 0.6 * info["pairs_total"], info
    # spheres ignore isenabled (Q4) unless the quirk is switched off: both policies agree with the dense path
    pc.enable_all()
    dense2, _ = R.score_counts(pc, cands, -1, params)
    got2, _ = R.score_counts_culled(pc, cands, params)
    np.testing.assert_array_equal(got2, dense2)
    assert (dense2 >= dense).all() and dense2.sum() > dense.sum()


@pytest.mark.parametrize("inline", ["0", "1"], ids=["queued", "inline"])
def test_culled_counts_on_adversarial_candidates(R, monkeypatch, inline):
    """non-unit axes and normals, wide / flat / needle cones, NaN / Inf / zero-axis candidates, points on axes"""
    from ransac_jl_b200 import _lib
    from ransac_jl_b200.shapes import from_cand
    from tests.helpers import adversarial_case

    monkeypatch.setenv("RSC_CULL_INLINE", inline)
    oshapes, P, N = adversarial_case()
    cands = []
    for sh in oshapes:
        c = _lib.rsc_cand(type=int(sh.kind), outwards=int(bool(sh.outwards)))
        for i, v in enumerate(sh.params7()):
            c.p[i] = float(v)
        cands.append(from_cand(c))
    pc = R.RANSACCloud(P.astype(np.float32), N.astype(np.float32), 1)
    pc.build_cells(6)
    params = R.ransacparameters()
    dense, _ = R.score_counts(pc, cands, -1, params)
    got, _ = R.score_counts_culled(pc, cands, params)
    np.testing.assert_array_equal(got, dense)


def test_culled_ragged_sizes_and_thresholds(R):
    from ransac_jl_b200 import scenes

    for n in (1, 511, 513, 5000):
        sc = scenes.scene_mixed(17 + n, max(n, 64), noise_frac=0.003, jitter_deg=1.0, outlier_frac=0.1, counts=(1, 1, 1, 1))
        V, N = sc.vertices[:n], sc.normals[:n]
        pc = R.RANSACCloud(V, N, 1)
        pc.build_cells(5)
        cands = scenes.perturbed_candidates(sc, 12, seed=5)
        for eps, alpha in ((0.3, np.deg2rad(5)), (1.5, np.deg2rad(20)), (0.01, np.deg2rad(1))):
            params = R.ransacparameters(plane={"eps": eps, "alpha": alpha}, sphere={"eps": eps, "alpha": alpha},
                                        cylinder={"eps": eps, "alpha": alpha}, cone={"eps": eps, "alpha": alpha})
            dense, _ = R.score_counts(pc, cands, -1, params)
            got, _ = R.score_counts_culled(pc, cands, params)
            np.testing.assert_array_equal(got, dense, err_msg=f"n={n} eps={eps}")


def test_culled_subset_counts_on_loop_like_candidates(R):
    """the path rsc_ransac_run takes for large batches: thousands of fitted candidates of interleaved types (as
    forcefitshapes! emits them) against a SUBSET copy (no build_cells: the subset is sorted on first use), with
    disabled points -- counts equal the dense path's; and again after the mask changes"""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(87, 60_000, noise_frac=0.002, jitter_deg=1.0, outlier_frac=0.05, counts=(3, 1, 1, 1))
    pc = R.RANSACCloud(sc.vertices, sc.normals, 4)
    params = R.ransacparameters(iteration={"tau": 600, "minsubsetN": 2048, "itermax": 40})
    cands, _, _ = R.sample_fit(pc, params, 3, 0, 32768)
    assert len(cands) > 1500 and len({type(c).__name__ for c in cands}) >= 3
    for sid in (0, 2):
        dense, _ = R.score_counts(pc, cands, sid, params)
        got, info = R.score_counts_culled(pc, cands, params, subsetID=sid)
        np.testing.assert_array_equal(got, dense, err_msg=f"subset {sid}")
        assert info["pairs_survived"] < info["pairs_total"]
    en = np.ones(pc.size, bool)
    en[np.random.default_rng(5).random(pc.size) < 0.4] = False
    pc.isenabled = en
    dense, _ = R.score_counts(pc, cands, 0, params)
    got, _ = R.score_counts_culled(pc, cands, params, subsetID=0)
    np.testing.assert_array_equal(got, dense)
    assert (dense > 0).sum() > 100


def test_culled_subset_view_follows_enable_all_and_new_coordinates(R):
    """the subset's Morton view is a cache: enable_all, isenabled assignments and a coordinate update must all reach it"""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(31, 40_000, noise_frac=0.002, jitter_deg=1.0, outlier_frac=0.1, counts=(2, 1, 1, 1))
    pc = R.RANSACCloud(sc.vertices, sc.normals, 2)
    params = R.ransacparameters()
    cands = scenes.perturbed_candidates(sc, 40, seed=5)

    def same():
        dense, _ = R.score_counts(pc, cands, 0, params)
        got, _ = R.score_counts_culled(pc, cands, params, subsetID=0)
        np.testing.assert_array_equal(got, dense)
        return int(dense.sum())

    full = same()
    en = np.ones(pc.size, bool)
    en[::3] = False
    pc.isenabled = en
    assert same() < full
    pc.enable_all()
    assert same() == full
    # other points under the same subset indices (rsc_cloud_update re-gathers the subset copies)
    from ransac_jl_b200._lib import lib

    V = np.ascontiguousarray(sc.vertices[::-1], dtype=np.float32)
    N = np.ascontiguousarray(sc.normals[::-1], dtype=np.float32)
    pc.ctx.check(lib.rsc_cloud_update(pc.handle, V.ctypes.data, N.ctypes.data, pc.size))
    assert same() > 0
