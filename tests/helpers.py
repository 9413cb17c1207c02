"""Shared helpers of the parity tests: conversions between the package's host types and the oracle's."""
import numpy as np

from oracle import ransac_oracle as O


def to_oracle_shape(sh) -> O.Shape:
    c = sh.to_cand()
    return O.shape_from_params7(c.type, bool(c.outwards), list(c.p))


def oracle_params(params) -> dict:
    """package params (shape_types are classes) -> oracle params (shape_types are kinds)"""
    import ransac_jl_b200 as R

    p = {k: dict(v) for k, v in params.items()}
    p["iteration"]["shape_types"] = [R.shapes.SHAPE_KIND[t] for t in params["iteration"]["shape_types"]]
    return p


def oracle_cloud(pc) -> O.Cloud:
    return O.Cloud(pc.vertices.astype(np.float64), pc.normals.astype(np.float64), [s.copy() for s in pc.subsets],
                   pc.isenabled.copy())


def oracle_mask(sh, pts, nrm, params, enabled=None, honour_enabled=True):
    m = O.compatibles(to_oracle_shape(sh), np.asarray(pts, np.float64), np.asarray(nrm, np.float64), params)
    if enabled is not None and honour_enabled:
        m = m & enabled
    return m


def near_threshold_report(sh, pts, nrm, params, idx):
    """For mismatching points: relative distance of the float64 margin terms to their thresholds
    (BASELINE.json: mismatches are tolerated only within 1e-6 relative of a threshold)."""
    from tests import fp32_model as M
    import math

    c = sh.to_cand()
    name = {0: "plane", 1: "sphere", 2: "cylinder", 3: "cone"}[c.type]
    eps, cosa = params[name]["eps"], math.cos(params[name]["alpha"])
    m = M.margin64(c.type, bool(c.outwards), list(c.p), np.asarray(pts)[idx], np.asarray(nrm)[idx], eps, cosa)
    return np.abs(m) / max(eps, 1e-300)
