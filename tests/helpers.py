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


def adversarial_case(seed: int = 12345, n: int = 4000):
    """Oracle shapes + float64 points/normals that stress the scorers: every shape type, wide/flat
    cones, non-unit axes and normals, huge and tiny radii, points on axes / at centres, zero normals,
    NaN / Inf / zero-axis candidates.  Coordinates are float32-representable (the device stores float32)."""
    import math

    rng = np.random.default_rng(seed)
    P = rng.normal(size=(n, 3)) * rng.choice([0.1, 1.0, 30.0], size=(n, 1))
    N = rng.normal(size=(n, 3))
    N /= np.linalg.norm(N, axis=1, keepdims=True)
    N[:50] *= rng.uniform(0.2, 3.0, (50, 1))  # non-unit normals (Q15)
    N[50:60] = 0.0                            # zero normals
    shapes = []
    for i in range(160):
        kind = i % 4
        a = rng.normal(size=3)
        if i % 5:
            a /= np.linalg.norm(a)            # otherwise a non-unit axis / plane normal (Q18)
        q = P[rng.integers(n)] + rng.normal(size=3) * rng.choice([0.0, 0.05, 1.0])
        outw = bool(rng.integers(2))
        if kind == 0:
            p7 = [*q, *a, 0.0]
        elif kind == 1:
            p7 = [*q, float(rng.choice([1e-3, 0.5, 3.0, 200.0])), 0, 0, 0]
        elif kind == 2:
            p7 = [*a, *q, float(rng.choice([1e-3, 0.5, 3.0, 200.0]))]
        else:
            p7 = [*q, *a, math.radians(float(rng.choice([1.0, 20.0, 90.0, 121.0, 179.0, 180.0, 200.0, 355.0])))]
        shapes.append(O.shape_from_params7(kind, outw, p7))
    # points exactly at a centre / on an axis (Q17), candidates with NaN / Inf parameters
    P[100] = shapes[1].params7()[0:3]
    P[101] = np.asarray(shapes[3].params7()[0:3]) + 2.0 * np.asarray(shapes[3].params7()[3:6])
    shapes.append(O.shape_from_params7(1, True, [np.nan, 0, 0, 1.0, 0, 0, 0]))
    shapes.append(O.shape_from_params7(3, True, [0, 0, 0, 0, 0, 1.0, np.inf]))
    shapes.append(O.shape_from_params7(2, False, [0, 0, 0, 1.0, 2.0, 3.0, 1.0]))  # zero axis
    P = P.astype(np.float32).astype(np.float64)
    N = N.astype(np.float32).astype(np.float64)
    return shapes, P, N


def degenerate_sets(seed: int = 777):
    """Minimal sets (S, 3, 3) points + normals that stress the fits: collinear points (Q3), coincident
    points, parallel / anti-parallel / zero normals, near-singular normal triples for the cone's rank
    test, tiny and huge scales, plus ordinary sets from a sphere, a cylinder and a cone for contrast."""
    rng = np.random.default_rng(seed)
    Ps, Ns = [], []

    def add(p, n):
        Ps.append(np.asarray(p, float)), Ns.append(np.asarray(n, float))

    e = np.eye(3)
    for s in (1e-3, 1.0, 1e3):
        add([[0, 0, 0], [s, 0, 0], [2 * s, 0, 0]], [e[2], e[2], e[2]])                # collinear, equal normals
        add([[0, 0, 0], [0, 0, 0], [s, s, 0]], [e[2], e[2], e[2]])                    # coincident points
        add([[0, 0, 0], [s, 0, 0], [0, s, 0]], [e[2], e[2], e[2]])                    # a clean plane
        add([[0, 0, 0], [s, 0, 0], [0, s, 0]], [e[2], -e[2], e[2]])                   # anti-parallel normal
        add([[0, 0, 0], [s, 0, 0], [0, s, 0]], [e[2], e[2], [0, 0, 0]])               # zero normal
        add([[s, 0, 0], [0, s, 0], [0, 0, s]], [e[0], e[1], e[2]])                    # sphere of radius s
        add([[s, 0, 0], [0, s, 0], [0, 0, s]], [-e[0], -e[1], -e[2]])                 # ... inward
        add([[s, 0, 0], [0, s, 0], [-s, 0, 5 * s]], [e[0], e[1], -e[0]])              # cylinder around z
        add([[s, 0, 0], [0, s, 0], [-s, 0, 5 * s]], [2 * e[0], 0.5 * e[1], -3 * e[0]])  # ... raw normals (Q8)
    for _ in range(40):  # cones: apex at a random place, random half angle, points on the surface
        apex, ax = rng.normal(size=3) * 5, rng.normal(size=3)
        ax /= np.linalg.norm(ax)
        half = rng.uniform(0.05, 1.4)
        u = np.cross(ax, rng.normal(size=3)); u /= np.linalg.norm(u)
        v = np.cross(ax, u)
        p, n = [], []
        for _k in range(3):
            t, ph = rng.uniform(0.5, 5.0), rng.uniform(0, 2 * np.pi)
            rad = np.cos(ph) * u + np.sin(ph) * v
            p.append(apex + t * (np.cos(half) * ax + np.sin(half) * rad))
            n.append(np.cos(half) * rad - np.sin(half) * ax)
        add(p, n)
    for _ in range(40):  # nearly coplanar normals: the cone's rank(r) == 3 decision sits near its tolerance
        a, b = rng.normal(size=3), rng.normal(size=3)
        a /= np.linalg.norm(a); b /= np.linalg.norm(b)
        c = 0.6 * a + 0.4 * b + rng.normal(size=3) * float(rng.choice([0.0, 1e-17, 1e-15, 1e-12, 1e-8, 1e-3]))
        add(rng.normal(size=(3, 3)) * 3, [a, b, c])
    for _ in range(60):  # random sets
        add(rng.normal(size=(3, 3)) * 10, rng.normal(size=(3, 3)))
    return np.stack(Ps), np.stack(Ns)
