"""Exchange fixtures that let an owner of Julia pin the oracle against the REAL RANSAC.jl.

Nothing here has been produced by Julia (there is none in this image).  This script writes
  * julia_inputs.json              -- a small seeded scene (float32-representable float64 coordinates), 24
                                      candidate shapes, 120 minimal sets, estimatescore arguments and, for one
                                      whole loop, the index triple of every minimal set of every iteration
                                      (0-based; null = failed sample);
  * julia_expected_by_oracle.json  -- what oracle/ransac_oracle.py (the float64 restatement of the
                                      reference's source) answers on those inputs.
`julia julia/make_golden.jl` runs the real package on julia_inputs.json and writes
tests/golden/julia_reference.json in the SAME format; tests/test_julia_golden.py compares the two
whenever that file exists (and checks this script's outputs against the oracles in any case).

Run from the repo root:  python tests/golden/make_julia_inputs.py
"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
NAMES = ["plane", "sphere", "cylinder", "cone"]


def shape_json(sh):
    p = [float(x) for x in sh.params7()]
    if sh.kind == 0:
        return {"type": "plane", "point": p[0:3], "normal": p[3:6]}
    if sh.kind == 1:
        return {"type": "sphere", "center": p[0:3], "radius": p[3], "outwards": bool(sh.outwards)}
    if sh.kind == 2:
        return {"type": "cylinder", "axis": p[0:3], "center": p[3:6], "radius": p[6], "outwards": bool(sh.outwards)}
    return {"type": "cone", "apex": p[0:3], "axis": p[3:6], "opang": p[6], "outwards": bool(sh.outwards)}


def shape_from_json(d):
    from oracle import ransac_oracle as O

    t = NAMES.index(d["type"])
    if t == 0:
        return O.shape_from_params7(0, True, [*d["point"], *d["normal"], 0.0])
    if t == 1:
        return O.shape_from_params7(1, d["outwards"], [*d["center"], d["radius"], 0, 0, 0])
    if t == 2:
        return O.shape_from_params7(2, d["outwards"], [*d["axis"], *d["center"], d["radius"]])
    return O.shape_from_params7(3, d["outwards"], [*d["apex"], *d["axis"], d["opang"]])


def oracle_params_from_json(p):
    op = {k: dict(v) for k, v in p.items()}
    op["iteration"]["shape_types"] = [NAMES.index(t) for t in p["iteration"]["shape_types"]]
    return op


def expected_by_oracle(inp):
    """the oracle's answers in the exchange format (shared by this script and tests/test_julia_golden.py)"""
    from oracle import ransac_oracle as O

    op = oracle_params_from_json(inp["params"])
    P, N = np.array(inp["points"], float), np.array(inp["normals"], float)
    subsets = [np.array(s, np.int64) for s in inp["subsets"]]
    cands = [shape_from_json(d) for d in inp["candidates"]]
    out = {"format": 1, "compatibles": [], "scorecandidate": [], "refit": [], "fits": [], "estimatescore": []}
    en = np.ones(len(P), bool)
    en[np.array(inp["disabled"], np.int64)] = False
    pc = O.Cloud(P, N, [s.copy() for s in subsets], en.copy())
    for sh in cands:
        out["compatibles"].append(np.flatnonzero(O.compatibles(sh, P, N, op)).tolist())
        ci, ip = O.scorecandidate(pc, sh, 0, op)
        out["scorecandidate"].append({"E": ci.E, "inpoints": ip.tolist()})
        out["refit"].append(O.refit(sh, pc, op).tolist())
    for sd in inp["minimal_sets"]:
        row = []
        for t in op["iteration"]["shape_types"]:
            sh = O.FIT[t](P[sd], N[sd], op)
            row.append(None if sh is None else shape_json(sh))
        out["fits"].append(row)
    for a in inp["estimatescore"]:
        ci = O.estimatescore(*a)
        out["estimatescore"].append([ci.min, ci.max, ci.E])
    sets = inp["loop"]["sets"]
    pc2 = O.Cloud(P, N, [s.copy() for s in subsets])
    tr = O.RansacTrace()
    ex = O.ransac(pc2, op, True, minimal_sets=lambda k, i: (None if k > len(sets) or sets[k - 1][i] is None else np.array(sets[k - 1][i])),
                  trace=tr)
    out["loop"] = {"iterations": tr.iterations, "extracted_at": tr.extracted_at,
                   "extracted": [dict(shape_json(e.shape), inpoints=e.inpoints.tolist()) for e in ex],
                   "isenabled": pc2.isenabled.astype(int).tolist()}
    return out


def main():
    from oracle import ransac_oracle as O
    from ransac_jl_b200 import scenes
    from tests.helpers import to_oracle_shape

    sc = scenes.scene_mixed(4242, 1600, noise_frac=0.002, jitter_deg=1.0, outlier_frac=0.1, counts=(2, 1, 1, 1))
    P, N = sc.vertices.astype(np.float64), sc.normals.astype(np.float64)
    rng = np.random.default_rng(99)
    perm = rng.permutation(len(P))
    subsets = [perm[:800], perm[800:]]
    params = {
        "iteration": {"drawN": 3, "minsubsetN": 40, "prob_det": 0.9, "tau": 60, "itermax": 50, "extract_s": "nofminset",
                      "terminate_s": "nofminset", "shape_types": ["plane", "cone", "cylinder", "sphere"]},
        "common": {"collin_threshold": 0.2, "parallelthrdeg": 1.0},
        "plane": {"eps": 0.3, "alpha": math.radians(5)}, "sphere": {"eps": 0.3, "alpha": math.radians(5), "sphere_par": 0.02},
        "cylinder": {"eps": 0.3, "alpha": math.radians(5)}, "cone": {"eps": 0.3, "alpha": math.radians(5), "minconeopang": math.radians(2)},
    }
    op = oracle_params_from_json(params)
    cands = [to_oracle_shape(p.shape) for p in sc.primitives] + [to_oracle_shape(s) for s in scenes.perturbed_candidates(sc, 4, seed=3)]
    cands += [O.shape_from_params7(3, True, [0, 0, 0, 0, 0, 1.0, math.radians(150)]),   # wide cone
              O.shape_from_params7(3, False, [1, 2, 3, 0.6, 0, 0.8, math.radians(200)]),  # beyond 180 degrees
              O.shape_from_params7(2, True, [0, 0, 2.0, 5, 5, 0, 3.0])]                  # non-unit axis (Q18)
    # minimal sets: same-primitive triples (so that fits succeed) and random triples
    labels = np.asarray(sc.labels)
    msets = []
    for j in range(len(sc.primitives)):
        idx = np.flatnonzero(labels == j)
        for _ in range(16):
            msets.append(rng.choice(idx, 3, replace=False).tolist())
    while len(msets) < 120:
        msets.append(rng.choice(len(P), 3, replace=False).tolist())
    disabled = np.sort(rng.choice(len(P), 200, replace=False)).tolist()
    # the loop: record the Philox sampler's index triples while the oracle's loop runs
    it = params["iteration"]
    pc = O.Cloud(P, N, [s.copy() for s in subsets])
    rec = []

    def ms(k, i):
        while len(rec) < k:
            rec.append([None] * it["minsubsetN"])
        ok, _, sd = O.sample_minimal_set(pc, 3, O.SetStream(777, (k - 1) * it["minsubsetN"] + i))
        rec[k - 1][i] = [int(x) for x in sd] if ok else None
        return sd if ok else None

    O.ransac(pc, op, True, minimal_sets=ms)
    inp = {"format": 1, "note": "indices are 0-based; coordinates are float32-representable float64",
           "params": params, "points": P.tolist(), "normals": N.tolist(), "subsets": [s.tolist() for s in subsets],
           "disabled": disabled, "candidates": [shape_json(c) for c in cands], "minimal_sets": msets,
           "estimatescore": [[800, 1600, 0], [800, 1600, 1], [800, 1600, 417], [800, 1600, 800], [312500, 10000000, 25000],
                             [1562500, 100000000, 900000]],
           "loop": {"sets": rec}}
    json.dump(inp, open(os.path.join(HERE, "julia_inputs.json"), "w"))
    json.dump(expected_by_oracle(inp), open(os.path.join(HERE, "julia_expected_by_oracle.json"), "w"))
    print("wrote julia_inputs.json,", len(P), "points,", len(cands), "candidates,", len(msets), "sets,", len(rec), "loop iterations")


if __name__ == "__main__":
    main()
