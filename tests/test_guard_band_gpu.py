"""On-device audit of the FP32 guard band: the margins the B200 really computes
(packed FFMA2 forms, MUFU.RSQ, ptxas' contraction) against the float64 margin of the same closed
forms, over 10^8 (candidate, point) pairs incl. points far from the origin, non-unit normals, tiny and
huge radii, narrow / wide / flat cones.  The kernel trusts an FP32 decision only if |m32| > band, so
the claim "every decision is the float64 one" needs |m32 - m64| < band for every pair; the test asks
for band / 2 and reports the worst ratio per shape type."""
import ctypes as C
import math

import numpy as np
import pytest

from tests import fp32_model as M

pytestmark = pytest.mark.gpu


def _cands(rng, sc, n_per_type):
    import ransac_jl_b200 as R
    from ransac_jl_b200 import scenes

    out = scenes.perturbed_candidates(sc, n_per_type, seed=int(rng.integers(1 << 30)), pos=0.5, ang_deg=3.0, rel=0.05)
    # the hard ones: tiny / huge radii, far centres, narrow, wide (> 120 deg) and flat (> 180 deg) cones, non-unit plane normal
    def u():
        v = rng.normal(size=3)
        return v / np.linalg.norm(v)

    far = lambda s: rng.uniform(-1, 1, 3) * s
    for R_ in (1e-2, 0.5, 50.0, 500.0):
        out.append(R.FittedSphere(far(100), R_, bool(rng.integers(2))))
        a = u()
        c = far(100)
        out.append(R.FittedCylinder(a, c - a * float(a @ c), R_, bool(rng.integers(2))))
    for deg in (2.5, 20.0, 89.0, 119.0, 121.0, 150.0, 179.0, 181.0, 240.0, 330.0):
        out.append(R.FittedCone(far(80), u(), math.radians(deg), bool(rng.integers(2))))
    out.append(R.FittedPlane(far(100), 3.0 * u()))
    out.append(R.FittedPlane(far(1000), u()))
    return out


@pytest.mark.parametrize("offset", [0.0, 2000.0], ids=["around_origin", "far_from_origin"])
def test_fp32_margin_error_is_inside_half_the_band(offset):
    import ransac_jl_b200 as R
    from ransac_jl_b200 import scenes
    from ransac_jl_b200._lib import lib

    rng = np.random.default_rng(11)
    npts = 1 << 20
    sc = scenes.scene_mixed(21, npts, noise_frac=0.01, jitter_deg=3.0, outlier_frac=0.3)
    V = (sc.vertices.astype(np.float64) + offset).astype(np.float32)
    Nn = sc.normals.copy()
    Nn[: npts // 8] *= rng.uniform(0.3, 2.5, (npts // 8, 1)).astype(np.float32)  # non-unit normals (Q15)
    sc = scenes.Scene(V, Nn, sc.labels, [scenes.Primitive(p.kind, p.shape, p.size, p.area, p.extra) for p in sc.primitives])
    pc = R.RANSACCloud(V, Nn, [np.zeros(0, np.int64)])
    cands = _cands(rng, scenes.Scene(sc.vertices - np.float32(offset), Nn, sc.labels, sc.primitives), 12)
    if offset:  # move the candidates with the cloud
        moved = []
        for s in cands:
            c = s.to_cand()
            p = list(c.p)
            if c.type in (0, 1, 3):
                p[0:3] = [x + offset for x in p[0:3]]
            else:
                ctr = np.array(p[3:6]) + offset
                a = np.array(p[0:3])
                p[3:6] = list(ctr - a * float(a @ ctr))
            moved.append(R.from_cand(R._lib.rsc_cand(c.type, c.outwards, (C.c_double * 7)(*p))))
        cands = moved
    Cn = len(cands)
    assert Cn * npts >= 7e7
    arr = R.pack_cands(cands)
    cp = R.to_c(R.ransacparameters())
    margins = np.zeros((Cn, npts), np.float32)
    bands = np.zeros(Cn, np.float32)
    cols = np.zeros(Cn, np.int32)
    diffs = C.c_int64()
    pc.ctx.check(lib.rsc_debug_margins(pc.handle, C.byref(cp), arr, Cn, 0, npts, margins.ctypes.data, bands.ctypes.data,
                                       cols.ctypes.data, C.byref(diffs)))
    assert diffs.value == 0, "packed (FFMA2) and scalar evaluation differ in some bit"
    pmax = float(np.sqrt((V.astype(np.float64) ** 2).sum(1).max()))
    nmax = float(np.sqrt((Nn.astype(np.float64) ** 2).sum(1).max()))
    P64, N64 = V.astype(np.float64), Nn.astype(np.float64)
    eps, cosa = 0.3, math.cos(math.radians(5))
    worst = {}
    for i, s in enumerate(cands):
        c = s.to_cand()
        _, band, scale = M.record(c.type, bool(c.outwards), list(c.p), pmax, nmax, eps, cosa)
        assert bands[i] == pytest.approx(band, rel=1e-5), "the device's band is not the documented kappa * 2^-24 * L"
        m64 = M.margin64(c.type, bool(c.outwards), list(c.p), P64, N64, eps, cosa) * scale
        m32 = margins[i].astype(np.float64)
        ok = np.isfinite(m64) & np.isfinite(m32)  # NaN margins (point on an axis / at a centre) go to FP64 by construction
        ratio = float(np.abs(m32[ok] - m64[ok]).max() / band)
        key = int(cols[i])
        worst[key] = max(worst.get(key, 0.0), ratio)
        assert ratio < 0.5, (i, c.type, int(cols[i]), list(c.p), ratio)
    print(f"offset {offset}: {Cn} candidates x {npts} points = {Cn * npts:.3g} pairs; worst |m32 - m64| / band by column type: "
          + ", ".join(f"{k}: {v:.3f}" for k, v in sorted(worst.items())))
    pc.close()
