"""The plug-in boundary of the reference (docs/src/newprimitive.md:12-18): a user-defined shape brings its
own fit / scorecandidate / refit and runs through the host loop, built-in shapes keep using the device
kernels.  For built-in shapes alone the host loop (per-call C ABI: rsc_sample_fit, rsc_score,
rsc_refit_extract) and the device loop (rsc_ransac_run) must give identical results."""
import math
from dataclasses import dataclass

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import ransac_jl_b200 as R

    return R


def test_host_loop_equals_device_loop(R):
    from ransac_jl_b200 import iterations as IT
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(86, 30_000, noise_frac=0.002, jitter_deg=1.0, outlier_frac=0.1, counts=(2, 1, 1, 1))
    params = R.ransacparameters(iteration={"tau": 300, "minsubsetN": 96, "itermax": 40})
    pc = R.RANSACCloud(sc.vertices, sc.normals, 4)
    pc.enable_all()
    dev, _ = IT._ransac_device(pc, params, 21)
    en_dev = pc.isenabled.copy()
    pc.enable_all()
    host, _ = IT._ransac_host(pc, params, 21)
    assert len(dev) == len(host) >= 3
    for a, b in zip(dev, host):
        assert type(a.shape) is type(b.shape)
        np.testing.assert_allclose(np.array(a.shape.to_cand().p[:7]), np.array(b.shape.to_cand().p[:7]), rtol=1e-12, atol=0)
        np.testing.assert_array_equal(a.inpoints, b.inpoints)
    np.testing.assert_array_equal(pc.isenabled, en_dev)


def test_user_defined_shape_plugs_in(R):
    """a z-slab "shape" (all points with |z - z0| < eps and a normal within alpha of +-z) written in NumPy,
    registered next to the built-in sphere: both are found, each by its own code path"""
    from ransac_jl_b200 import scenes

    @dataclass
    class FittedSlab(R.FittedShape):
        z0: float

        @staticmethod
        def defaultshapeparameters():  # fitting.jl:15
            return {"slab": {"eps": 0.3, "alpha": math.radians(5)}}

        @staticmethod
        def fit(p, n, pc, params):
            z = np.asarray(p, float)[:, 2]
            nz = np.abs(np.asarray(n, float)[:, 2])
            if z.max() - z.min() < 0.1 and (nz > math.cos(math.radians(5))).all():
                return FittedSlab(float(z.mean()))
            return None

        def _mask(self, v, n):
            return (np.abs(v[:, 2] - self.z0) < 0.3) & (np.abs(n[:, 2]) > math.cos(math.radians(5)))

        def scorecandidate(self, pc, subsetID, params):
            sub = pc.subsets[subsetID]
            m = self._mask(pc.vertices[sub], pc.normals[sub]) & pc.isenabled[sub]
            return R.estimatescore(len(sub), pc.size, int(m.sum())), sub[m]

        def refit(self, pc, params):
            m = self._mask(pc.vertices, pc.normals) & pc.isenabled
            return R.ExtractedShape(self, np.flatnonzero(m))

    rng = np.random.default_rng(9)
    n_slab, n_sph, n_out = 6000, 5000, 2000
    slab = np.c_[rng.uniform(-20, 20, (n_slab, 2)), np.full(n_slab, 7.0) + rng.normal(0, 0.02, n_slab)]
    slab_n = np.tile([0.0, 0.0, 1.0], (n_slab, 1))
    d = rng.normal(size=(n_sph, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    sph, sph_n = np.array([30.0, 0, -10]) + 6.0 * d, d
    out = rng.uniform(-40, 40, (n_out, 3))
    out_n = rng.normal(size=(n_out, 3))
    out_n /= np.linalg.norm(out_n, axis=1, keepdims=True)
    V = np.r_[slab, sph, out].astype(np.float32)
    N = np.r_[slab_n, sph_n, out_n].astype(np.float32)
    perm = rng.permutation(len(V))
    V, N = V[perm], N[perm]
    pc = R.RANSACCloud(V, N, 2)
    params = R.ransacparameters([R.FittedSphere, FittedSlab], iteration={"tau": 500, "minsubsetN": 64, "itermax": 60})
    extracted, _ = R.ransac(pc, params, True, seed=3)
    kinds = [type(e.shape).__name__ for e in extracted]
    assert "FittedSlab" in kinds and "FittedSphere" in kinds, kinds
    slab_ex = next(e for e in extracted if isinstance(e.shape, FittedSlab))
    sph_ex = next(e for e in extracted if isinstance(e.shape, R.FittedSphere))
    assert abs(slab_ex.shape.z0 - 7.0) < 0.05 and len(slab_ex.inpoints) > 0.95 * n_slab
    assert abs(sph_ex.shape.radius - 6.0) < 0.05 and len(sph_ex.inpoints) > 0.95 * n_sph
    taken = np.concatenate([e.inpoints for e in extracted])
    assert len(np.unique(taken)) == len(taken)
    assert pc.count_enabled() == pc.size - len(taken)
