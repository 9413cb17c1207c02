"""The culling criterion of rsc_score_culled (csrc/rsc_cull.cu::cull_far), modelled in NumPy float32:
a (candidate, Morton tile) pair may only be skipped if the oracle finds no compatible point in the tile -- at
each of the three tile sizes the kernel tests (4096-point blocks in its pre-pass, 512-point groups, 128-point
tiles).  CPU test of the math (Lipschitz bounds, scaled cone records, non-unit cylinder axes, far
records of NaN/Inf candidates); the kernel itself is compared with the dense path on the GPU."""
import math

import numpy as np

from oracle import ransac_oracle as O
from tests import fp32_model as M

f32 = np.float32
TILES = (128, 512, 4096)


def morton_tiles(V, levels=9, TILE=512):
    lo, hi = V.min(0), V.max(0)
    D = 1 << (levels - 1)
    w = np.where(hi > lo, hi - lo, 1.0)
    q = np.minimum(((V - lo) / w * D).astype(np.int64), D - 1)
    order = np.argsort(O.morton3(q, D), kind="stable")
    tiles = []
    for s in range(0, len(V), TILE):
        idx = order[s : s + TILE]
        P = V[idx].astype(f32)
        c = (f32(0.5) * (P.min(0) + P.max(0))).astype(f32)
        d = (P - c).astype(f32)
        r2 = (d * d).sum(1).astype(f32).max()
        r = f32(np.sqrt(r2)) * f32(1.00001) + f32(1e-30)
        tiles.append((idx, c, r))
    return tiles


def _fma(a, b, c):
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def cull_far32(kind, rec, band, c, rt, eps):
    """float32 model of cull_far for one candidate record and one tile sphere"""
    r = [f32(x) for x in rec]
    eps, rt = f32(eps), f32(rt)
    cx, cy, cz = f32(c[0]), f32(c[1]), f32(c[2])
    with np.errstate(all="ignore"):
        if kind == 0:
            d = abs(_fma(r[0], cx, _fma(r[1], cy, _fma(r[2], cz, r[3]))))
            lim = eps + rt
        else:
            vx, vy, vz = _fma(r[0], cx, r[1]), _fma(r[0], cy, r[2]), _fma(r[0], cz, r[3])
            if kind == 1:
                d = abs(f32(np.sqrt(_fma(vx, vx, _fma(vy, vy, f32(vz * vz))))) + r[4])
                lim = eps + rt
            else:
                h = _fma(r[4], vx, _fma(r[5], vy, f32(r[6] * vz)))
                wx, wy, wz = _fma(-r[4], h, vx), _fma(-r[5], h, vy), _fma(-r[6], h, vz)
                rho = f32(np.sqrt(_fma(wx, wx, _fma(wy, wy, f32(wz * wz)))))
                if kind == 2:
                    a2 = _fma(r[4], r[4], _fma(r[5], r[5], f32(r[6] * r[6])))
                    d = abs(rho + r[7])
                    lim = eps + rt * max(f32(1.0), abs(a2 - f32(1.0)) * f32(1.0001))
                elif len(r) == 11:
                    d = abs(_fma(-rho, r[7], h))
                    lim = -r[8] * (f32(1.0) + rt / eps)
                else:
                    d = abs(_fma(h, r[7], -rho))
                    lim = -r[8] * (f32(1.0) + rt / eps)
        return bool(f32(d) > f32(lim) * f32(1.0001) + f32(8.0) * f32(band))


def check_no_false_culls(shapes, V, N, params, levels, tile=512):
    tiles = morton_tiles(V, levels, tile)
    pmax = float(np.sqrt((V.astype(f32).astype(np.float64) ** 2).sum(1).max()))
    nmax = float(np.sqrt((N.astype(f32).astype(np.float64) ** 2).sum(1).max()))
    culled = total = 0
    for sh in shapes:
        p = sh.params7()
        name = O.SHAPE_NAMES[sh.kind]
        eps, cosa = params[name]["eps"], math.cos(params[name]["alpha"])
        if not np.isfinite(p).all():
            continue  # "far" records: the kernel may cull or evaluate them, the count is 0 either way
        if sh.kind == O.PLANE and not np.linalg.norm(p[3:6]) > 0:
            continue
        if sh.kind == O.CONE:
            if not np.linalg.norm(p[3:6]) > 0:
                continue
            if math.cos(p[6] / 2) < 0.5 and not math.sin(p[6] / 2) >= 1 / 16:
                continue  # needle: infinite band, never culled
        rec, band, _ = M.record(sh.kind, sh.outwards, p, pmax, nmax, eps, cosa)
        comp = O.compatibles(sh, V, N, params)
        for idx, c, rt in tiles:
            total += 1
            if cull_far32(sh.kind, rec, band, c, rt, eps):
                culled += 1
                assert not comp[idx].any(), (name, p, c, rt)
    return culled, total


def test_no_false_culls_on_a_noisy_scene():
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(91, 40_000, noise_frac=0.004, jitter_deg=1.5, outlier_frac=0.2, counts=(2, 1, 1, 1))
    V = sc.vertices.astype(np.float64)
    N = sc.normals.astype(np.float64)
    from tests.helpers import to_oracle_shape

    shapes = [to_oracle_shape(s) for s in scenes.perturbed_candidates(sc, 12, seed=7)]
    P = O.default_parameters()
    for tile in TILES:
        culled, total = check_no_false_culls(shapes, V, N, P, 8, tile)
        assert culled > (0.3 if tile <= 512 else 0.1) * total, (tile, culled, total)  # and the test culls something
    P2 = O.ransacparameters(P, plane={"eps": 2.0}, sphere={"eps": 2.0}, cylinder={"eps": 2.0}, cone={"eps": 2.0})
    check_no_false_culls(shapes, V, N, P2, 8)


def test_no_false_culls_on_adversarial_candidates():
    from tests.helpers import adversarial_case

    shapes, V, N = adversarial_case()
    for tile in TILES:
        culled, total = check_no_false_culls(shapes, V, N, O.default_parameters(), 6, tile)
        assert culled > 0 or tile == 4096
