"""Generates the committed fixtures of tests/golden/.

PROVENANCE -- read before trusting these files.  The reference (cserteGT3/RANSAC.jl v0.6.0) is a Julia
package and no Julia toolchain exists in this image, so the reference itself cannot produce vectors.
  * reference_known_answers.json is a TRANSCRIPTION of the known-answer tests the reference ships
    (test/dummyspheretest.jl, test/confidenceintervals.jl, test/utilitytests.jl; file:line per case).
    It is written by hand below, not generated.
  * oracle_small_scene.npz holds OUTPUTS OF THE ORACLE (oracle/ransac_oracle.py, the float64
    restatement of the reference's source) on a small seeded scene.  They pin the CUDA path, the C
    oracle and the NumPy oracle to each other across commits; they are regression pins, not
    reference outputs ("parity unpinned" in DESIGN.md section 5 still applies to them).

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

HERE = os.path.dirname(os.path.abspath(__file__))

KNOWN = {
    "source": "cserteGT3/RANSAC.jl v0.6.0, test/ (transcribed; file:line per case)",
    "dummyspheretest": {
        "cite": "test/dummyspheretest.jl:8-49",
        "sphere_params": {"eps": 0.1, "alpha_deg": 10.0},
        "plane_params": {"alpha": math.pi / 2, "collin_threshold": 0.2},
        "normals": [[0, -1, 0.0], [0, 0, -1.0], [1, 0, 0.0], [0, 1, 0.0]],
        "cases": [
            {"cite": "test/dummyspheretest.jl:14-22", "name": "true sphere 1",
             "points": [[0, -1, 0.0], [0, 0, -1.0], [1, 0, 0.0], [0, 1, 0.0]],
             "sphere": {"accept": True, "center": [0, 0, 0], "radius": 1.0, "outwards": True}, "plane_accept": False},
            {"cite": "test/dummyspheretest.jl:24-35", "name": "true sphere 2",
             "points": [[0, -0.99, 0.0], [0, 0, -1.0], [1.01, 0, 0.0], [0, 1, 0.0]],
             "sphere": {"accept": True}, "sphere_eps_0.01": {"accept": False}, "plane_accept": False},
            {"cite": "test/dummyspheretest.jl:37-49", "name": "false sphere 1",
             "points": [[0, 1, 0.0], [0, 0, -1.0], [1, 0, 0.0], [0, 1, 0.0]],
             "sphere": {"accept": False}, "sphere_eps_10_alpha_pi_2": {"accept": False}, "plane_accept": False},
        ],
    },
    "confidenceintervals": {
        "cite": "test/confidenceintervals.jl:1-25",
        "ctor": {"min": 1.0, "max": 3, "E": 2.0, "reversed_throws": True},
        "notsoconfident": {"x": 153.9, "y": 9.7, "min": 9.7, "max": 153.9, "E": 81.8},
    },
    "defaults": {
        "cite": "test/utilitytests.jl:41-82, src/RANSAC.jl:94",
        "iteration": {"drawN": 3, "minsubsetN": 15, "prob_det": 0.9, "tau": 900, "itermax": 1000,
                      "extract_s": "nofminset", "terminate_s": "nofminset",
                      "shape_types": ["plane", "cone", "cylinder", "sphere"]},
        "common": {"collin_threshold": 0.2, "parallelthrdeg": 1.0},
        "plane": {"eps": 0.3, "alpha_deg": 5.0},
        "sphere": {"eps": 0.3, "alpha_deg": 5.0, "sphere_par": 0.02},
        "cylinder": {"eps": 0.3, "alpha_deg": 5.0},
        "cone": {"eps": 0.3, "alpha_deg": 5.0, "minconeopang_deg": 2.0},
    },
}


def main():
    with open(os.path.join(HERE, "reference_known_answers.json"), "w") as f:
        json.dump(KNOWN, f, indent=1)

    from oracle import ransac_oracle as O
    from ransac_jl_b200 import scenes
    from ransac_jl_b200.params import ransacparameters
    from tests.helpers import oracle_params, to_oracle_shape
    import ransac_jl_b200.shapes as SH

    sc = scenes.scene_mixed(2024, 3000, noise_frac=0.004, jitter_deg=1.5, outlier_frac=0.15, counts=(2, 1, 1, 1))
    P32, N32 = sc.vertices, sc.normals
    P, N = P32.astype(np.float64), N32.astype(np.float64)
    params = ransacparameters()
    op = oracle_params(params)
    cands = [p.shape for p in sc.primitives] + scenes.perturbed_candidates(sc, 4, seed=11)
    rng = np.random.default_rng(5)
    a = rng.normal(size=3); a /= np.linalg.norm(a)
    cands.append(SH.FittedCone(P[7] - 0.4 * a, a, math.radians(150.0), True))     # wide cone (column type kConeWide)
    # nearly flat cones lying in the first plane (what three-point cone fits on planar patches give)
    pl = sc.primitives[0].shape
    c0 = P[lab0 := np.flatnonzero(sc.labels == 0)].mean(0)
    for outw in (True, False):
        cands.append(SH.FittedCone(c0, np.asarray(pl.normal, float), math.radians(176.0), outw))
        cands.append(SH.FittedCone(c0, -np.asarray(pl.normal, float), math.radians(183.0), outw))
    enabled = rng.random(len(P)) > 0.2
    c7 = np.array([[float(x) for x in sh.to_cand().p[:7]] for sh in cands])
    ctype = np.array([sh.to_cand().type for sh in cands], np.int32)
    coutw = np.array([int(sh.to_cand().outwards) for sh in cands], np.int32)
    masks = np.stack([O.compatibles(to_oracle_shape(sh), P, N, op) for sh in cands])
    # minimal sets: 300 in-primitive triples + 100 random triples
    lab = sc.labels
    pools = [np.flatnonzero(lab == l) for l in range(lab.max() + 1)]
    sets = [rng.choice(pools[rng.integers(len(pools))], 3, replace=False) for _ in range(300)]
    sets += [rng.choice(len(P), 3, replace=False) for _ in range(100)]
    sets = np.stack(sets).astype(np.int64)
    fit_kind, fit_set, fit_outw, fit_p7 = [], [], [], []
    for si, idx in enumerate(sets):
        for sh in O.forcefit(P[idx], N[idx], op):
            fit_kind.append(sh.kind), fit_set.append(si), fit_outw.append(int(bool(sh.outwards))), fit_p7.append(sh.params7())
    # whole loop on the same cloud, 2 subsets from a fixed permutation
    perm = np.random.default_rng(77).permutation(len(P)).astype(np.int64)
    subsets = [perm[: len(P) // 2].copy(), perm[len(P) // 2:].copy()]
    it = {"tau": 150, "minsubsetN": 64, "itermax": 60}
    rp = ransacparameters(iteration=it)
    ex = O.ransac(O.Cloud(P32, N32, [s.copy() for s in subsets]), oracle_params(rp), True, seed=31)
    np.savez_compressed(
        os.path.join(HERE, "oracle_small_scene.npz"),
        vertices=P32, normals=N32, enabled=enabled,
        cand_type=ctype, cand_outwards=coutw, cand_p7=c7, masks=np.packbits(masks, axis=1, bitorder="little"),
        fit_sets=sets, fit_kind=np.array(fit_kind, np.int32), fit_set=np.array(fit_set, np.int32),
        fit_outwards=np.array(fit_outw, np.int32), fit_p7=np.array(fit_p7, np.float64),
        subset0=subsets[0], subset1=subsets[1], run_iteration=np.array([it["tau"], it["minsubsetN"], it["itermax"], 31]),
        run_kind=np.array([e.shape.kind for e in ex], np.int32), run_p7=np.array([e.shape.params7() for e in ex], np.float64),
        run_len=np.array([len(e.inpoints) for e in ex], np.int64),
        run_inpoints=np.concatenate([e.inpoints for e in ex]) if ex else np.zeros(0, np.int64),
    )
    print(f"{len(cands)} candidates, inlier counts {masks.sum(1).tolist()}; {len(fit_kind)} fitted candidates from {len(sets)} sets; "
          f"loop: {[O.SHAPE_NAMES[e.shape.kind] for e in ex]}")


if __name__ == "__main__":
    main()
