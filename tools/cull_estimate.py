"""CPU estimate for the round-2 lever named in DESIGN.md section 3: tile culling.

With the cloud in Morton order, a tile of T consecutive points has a small bounding sphere (centre c,
radius r).  Every distance function of the path (plane.jl:82-103, sphere.jl:163-166, cylinder.jl:209-214,
cone closed form h*sin - rho*cos) is 1-Lipschitz in the point, so |dist(c)| > eps + r proves that NO point
of the tile is compatible with the candidate: the (candidate, tile) pair can be skipped with counts and
masks unchanged.  This script measures, on the c3 workload generator at reduced N, which fraction of the
(candidate, point) evaluations survives that test (NumPy, oracle formulas; nothing here is shipped).

    python tools/cull_estimate.py --points 1048576 --tile 512
"""
import argparse
import json
import math
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ransac_oracle as O  # noqa: E402
from ransac_jl_b200 import scenes  # noqa: E402  (scene generators only: NumPy, no device needed)


def dist_to(shape, c):
    """signed distance of points c (m,3) to an oracle shape, float64"""
    if shape.kind == O.PLANE:
        n = shape.b / np.linalg.norm(shape.b)
        return (c - shape.a) @ n
    if shape.kind == O.SPHERE:
        return np.linalg.norm(c - shape.a, axis=1) - shape.s
    if shape.kind == O.CYLINDER:
        a = shape.a / np.linalg.norm(shape.a)
        v = c - shape.b
        w = v - (v @ a)[:, None] * a
        return np.linalg.norm(w, axis=1) - shape.s
    a = shape.b / np.linalg.norm(shape.b)
    v = c - shape.a
    h = v @ a
    rho = np.linalg.norm(v - h[:, None] * a, axis=1)
    return h * math.sin(shape.s / 2) - rho * math.cos(shape.s / 2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=1 << 20)
    ap.add_argument("--tile", type=int, default=512)
    ap.add_argument("--per-type", type=int, default=64)
    ap.add_argument("--levels", type=int, default=11)
    args = ap.parse_args()
    sc = scenes.scene_c3(args.points)
    V = sc.vertices.astype(np.float64)
    lo, hi = V.min(0), V.max(0)
    D = 1 << (args.levels - 1)
    q = np.minimum(((V - lo) / (hi - lo) * D).astype(np.int64), D - 1)
    order = np.argsort(O.morton3(q, D), kind="stable")
    Vs = V[order]
    nt = len(Vs) // args.tile
    T = Vs[: nt * args.tile].reshape(nt, args.tile, 3)
    c = (T.min(1) + T.max(1)) / 2
    r = np.linalg.norm(T - c[:, None, :], axis=2).max(1)
    cands = scenes.perturbed_candidates(sc, args.per_type, seed=7)
    eps = 0.3
    out = {}
    for kind, name in O.SHAPE_NAMES.items():
        fr = []
        for s in cands:
            cd = s.to_cand() if hasattr(s, "to_cand") else None
            sh = O.shape_from_params7(cd.type, bool(cd.outwards), list(cd.p)) if cd is not None else s
            if sh.kind != kind:
                continue
            d = dist_to(sh, c)
            fr.append(float((np.abs(d) <= eps * 1.001 + r).mean()))
        out[name] = round(float(np.mean(fr)), 4) if fr else None
    allf = [v for v in out.values() if v is not None]
    print(json.dumps({"points": args.points, "tile": args.tile, "tiles": nt, "median_tile_radius": round(float(np.median(r)), 3),
                      "surviving_fraction_by_type": out, "surviving_fraction_mean": round(float(np.mean(allf)), 4)}))


if __name__ == "__main__":
    main()
