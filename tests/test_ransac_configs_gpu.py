"""End-to-end parity of the device loop at the BASELINE.json sizes: config c2 (1 Mi points, 6 planes +
3 spheres + 2 cylinders + 2 cones, noise + 20 % outliers) and config c4 (10 M points, 200 primitives)
against the C oracle's loop (oracle/oracle.c::orc_ransac -- the reference's iterations.jl:35-162
restated; tests/test_c_oracle_loop.py pins it to the NumPy oracle) on the same Philox minimal sets:
same shapes in the same order, identical inlier index lists, identical final isenabled."""
import numpy as np
import pytest

from oracle import c_oracle
from tests.helpers import oracle_params

pytestmark = pytest.mark.gpu


def _check(R, sc, r, it, seed):
    params = R.ransacparameters(iteration=it)
    pc = R.RANSACCloud(sc.vertices, sc.normals, r)
    extracted, _ = R.ransac(pc, params, True, seed=seed)
    want, en, info = c_oracle.ransac(sc.vertices, sc.normals, pc.subsets[0], oracle_params(params), seed)
    assert len(extracted) == len(want), (len(extracted), len(want))
    for got, w in zip(extracted, want):
        c = got.shape.to_cand()
        assert c.type == w[0]
        if c.type != 0:
            assert bool(c.outwards) == w[1]
        np.testing.assert_allclose(np.array(c.p[:]), w[2], rtol=1e-5, atol=1e-5 * max(1.0, np.abs(w[2]).max()))
        np.testing.assert_array_equal(got.inpoints, w[3])
    np.testing.assert_array_equal(pc.isenabled, en)
    pc.close()
    return extracted, info


@pytest.mark.parametrize("loop_cull", ["1", "2"], ids=["default", "culled-scorer-for-every-batch"])
def test_c2_full_loop_matches_c_oracle(monkeypatch, loop_cull):
    """RSC_LOOP_CULL=2: every batch of new candidates is scored by the culled scorer on the subset's Morton view
    (by default only batches of >= 4e8 pairs are) -- same shapes, same lists"""
    import ransac_jl_b200 as R
    from ransac_jl_b200 import scenes

    monkeypatch.setenv("RSC_LOOP_CULL", loop_cull)
    sc = scenes.scene_c2()
    ex, info = _check(R, sc, 32, {"tau": len(sc.vertices) // 100, "minsubsetN": 4096, "itermax": 200}, 2024)
    assert len(ex) >= 13 and info["iterations"] == 200


def test_c4_full_loop_matches_c_oracle():
    import ransac_jl_b200 as R
    from ransac_jl_b200 import scenes

    sc = scenes.scene_cad()
    ex, info = _check(R, sc, 32, {"tau": len(sc.vertices) // 1000, "minsubsetN": 8192, "itermax": 400}, 2024)
    assert len(ex) >= 30 and info["iterations"] == 400
    print(f"c4: {len(ex)} shapes; C oracle loop {sum(info['seconds']):.1f} s on {info['threads']} threads")
