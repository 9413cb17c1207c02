"""The CUDA path against the committed fixtures of tests/golden/oracle_small_scene.npz -- no oracle
is imported here: counts, bitmasks, refit lists, fits and the whole loop must equal the stored answers."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def R():
    import ransac_jl_b200 as R

    return R


@pytest.fixture(scope="module")
def small():
    return np.load(os.path.join(GOLD, "oracle_small_scene.npz"))


def _shapes(R, small):
    from ransac_jl_b200 import _lib
    from ransac_jl_b200.shapes import from_cand

    out = []
    for t, o, p in zip(small["cand_type"], small["cand_outwards"], small["cand_p7"]):
        c = _lib.rsc_cand(type=int(t), outwards=int(o))
        for i in range(7):
            c.p[i] = float(p[i])
        out.append(from_cand(c))
    return out


def test_counts_masks_refit(R, small):
    P, N = small["vertices"], small["normals"]
    want = np.unpackbits(small["masks"], axis=1, bitorder="little")[:, : len(P)].astype(bool)
    en = small["enabled"].astype(bool)
    pc = R.RANSACCloud(P, N, 1)
    params = R.ransacparameters()
    cands = _shapes(R, small)
    counts, masks = R.score_counts(pc, cands, -1, params, want_masks=True)
    np.testing.assert_array_equal(counts, want.sum(1))
    for i in range(len(cands)):
        np.testing.assert_array_equal(R.unpack_mask(masks[i], pc.size), want[i])
    pc.isenabled = en
    counts, masks = R.score_counts(pc, cands, -1, params, want_masks=True)
    for i, sh in enumerate(cands):
        gated = want[i] if isinstance(sh, R.FittedSphere) else (want[i] & en)  # Q4: spheres ignore isenabled
        assert counts[i] == gated.sum()
        np.testing.assert_array_equal(R.unpack_mask(masks[i], pc.size), gated)
        np.testing.assert_array_equal(R.refit(sh, pc, params).inpoints, np.flatnonzero(want[i] & en))


def test_fits(R, small):
    pc = R.RANSACCloud(small["vertices"], small["normals"], 1)
    shapes, sets = R.fit_batch(pc, small["fit_sets"], R.ransacparameters())
    kinds = [sh.to_cand().type for sh in shapes]
    assert kinds == small["fit_kind"].tolist()
    assert list(sets) == small["fit_set"].tolist()
    for sh, k, o, p in zip(shapes, small["fit_kind"], small["fit_outwards"], small["fit_p7"]):
        c = sh.to_cand()
        if k != 0:
            assert int(c.outwards) == int(o)
        np.testing.assert_allclose(np.array(c.p[:7]), p, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(p).max()))


def test_loop(R, small):
    tau, msn, itmax, seed = (int(x) for x in small["run_iteration"])
    pc = R.RANSACCloud(small["vertices"], small["normals"], [small["subset0"], small["subset1"]])
    params = R.ransacparameters(iteration={"tau": tau, "minsubsetN": msn, "itermax": itmax})
    ex, _ = R.ransac(pc, params, True, seed=seed)
    assert [e.shape.to_cand().type for e in ex] == small["run_kind"].tolist()
    assert [len(e.inpoints) for e in ex] == small["run_len"].tolist()
    np.testing.assert_array_equal(np.concatenate([e.inpoints for e in ex]), small["run_inpoints"])
    for e, p in zip(ex, small["run_p7"]):
        np.testing.assert_allclose(np.array(e.shape.to_cand().p[:7]), p, rtol=1e-5, atol=1e-5 * max(1.0, np.abs(p).max()))
