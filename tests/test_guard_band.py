"""CPU check of the FP32 guard band: on seeded scenes, the float32 model of the kernel arithmetic
never strays from the float64 margin by more than half the band the kernel uses, so every pair the
kernel decides in FP32 (|margin| > band) is decided like float64."""
import math

import numpy as np

import ransac_jl_b200.scenes as S
from tests import fp32_model as M


def _cand_params(sh):
    c = sh.to_cand()
    return c.type, bool(c.outwards), list(c.p)


def test_fp32_error_within_band():
    sc = S.scene_mixed(11, 60_000)
    P, N = sc.vertices, sc.normals
    pmax = float(np.sqrt((P.astype(np.float64) ** 2).sum(1)).max())
    nmax = float(np.sqrt((N.astype(np.float64) ** 2).sum(1)).max())
    cands = S.perturbed_candidates(sc, 24, seed=5)
    eps, cosa = 0.3, math.cos(math.radians(5))
    worst = {}
    for sh in cands:
        kind, outw, p = _cand_params(sh)
        r, band, scale = M.record(kind, outw, p, pmax, nmax, eps, cosa)
        m32 = M.margin32(kind, r, P, N, eps, cosa).astype(np.float64)
        m64 = M.margin64(kind, outw, p, P, N, eps, cosa) * scale
        ok = np.isfinite(m32) & np.isfinite(m64)
        ratio = float((np.abs(m32 - m64)[ok] / band).max())
        worst[kind] = max(worst.get(kind, 0.0), ratio)
    print("max |m32-m64|/band per type:", worst)
    for kind, ratio in worst.items():
        assert ratio < 0.5, (kind, ratio)


def test_fp32_far_candidates_within_band():
    """candidates far outside the cloud (large |apex|, large R) keep the bound"""
    rng = np.random.default_rng(3)
    sc = S.scene_mixed(12, 20_000)
    P, N = sc.vertices, sc.normals
    pmax = float(np.sqrt((P.astype(np.float64) ** 2).sum(1)).max())
    eps, cosa = 0.3, math.cos(math.radians(5))
    for kind in range(4):
        for _ in range(12):
            a = rng.normal(size=3); a /= np.linalg.norm(a)
            far = rng.normal(size=3) * 2000.0
            if kind == 0:
                p = [*far, *a, 0]
            elif kind == 1:
                p = [*far, float(np.linalg.norm(far)), 0, 0, 0]
            elif kind == 2:
                c = far - a * float(a @ far)
                p = [*a, *c, float(np.linalg.norm(c))]
            else:
                p = [*far, *(-far / np.linalg.norm(far) * 0.9 + a * 0.1), 0.3]
                ax = np.array(p[3:6]); p[3:6] = list(ax / np.linalg.norm(ax))
            r, band, scale = M.record(kind, True, p, pmax, 1.0, eps, cosa)
            m32 = M.margin32(kind, r, P, N, eps, cosa).astype(np.float64)
            m64 = M.margin64(kind, True, p, P, N, eps, cosa) * scale
            ok = np.isfinite(m32) & np.isfinite(m64)
            assert (np.abs(m32 - m64)[ok] / band).max() < 0.5, kind


def test_fp32_wide_cones_within_band():
    """cones wider than 120 degrees use the 1/sin(opang/2)-scaled form (column type kConeWide)"""
    rng = np.random.default_rng(4)
    sc = S.scene_mixed(13, 30_000)
    P, N = sc.vertices, sc.normals
    pmax = float(np.sqrt((P.astype(np.float64) ** 2).sum(1)).max())
    eps, cosa = 0.3, math.cos(math.radians(5))
    worst = 0.0
    for deg in (100.0, 119.0, 121.0, 150.0, 175.0, 179.9, 180.0, 185.0, 270.0, 340.0):
        for outw in (True, False):
            a = rng.normal(size=3); a /= np.linalg.norm(a)
            apex = P[rng.integers(len(P))].astype(np.float64) - a * rng.uniform(0, 3)
            p = [*apex, *a, math.radians(deg)]
            r, band, scale = M.record(3, outw, p, pmax, 1.0, eps, cosa)
            assert (len(r) == 11) == (deg > 120.0)
            m32 = M.margin32(3, r, P, N, eps, cosa).astype(np.float64)
            m64 = M.margin64(3, outw, p, P, N, eps, cosa) * scale
            ok = np.isfinite(m32) & np.isfinite(m64)
            worst = max(worst, float((np.abs(m32 - m64)[ok] / band).max()))
    print("wide cones: max |m32-m64|/band", worst)
    assert worst < 0.5
