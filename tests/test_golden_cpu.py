"""The oracles against the committed fixtures of tests/golden/ (CPU only).

reference_known_answers.json transcribes the reference's own known-answer tests; oracle_small_scene.npz
holds oracle outputs that pin NumPy oracle, C oracle and CUDA path to each other (see make_golden.py)."""
import json
import math
import os

import numpy as np
import pytest

from oracle import ransac_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def known():
    return json.load(open(os.path.join(GOLD, "reference_known_answers.json")))


@pytest.fixture(scope="module")
def small():
    return np.load(os.path.join(GOLD, "oracle_small_scene.npz"))


def test_reference_sphere_plane_cases(known):
    d = known["dummyspheretest"]
    base = O.ransacparameters(O.ransacparameters(), sphere={"eps": d["sphere_params"]["eps"],
                                                            "alpha": math.radians(d["sphere_params"]["alpha_deg"])})
    prp = O.ransacparameters(base, plane={"alpha": d["plane_params"]["alpha"]},
                             common={"collin_threshold": d["plane_params"]["collin_threshold"]})
    tn = [tuple(x) for x in d["normals"]]
    for case in d["cases"]:
        tv = [tuple(x) for x in case["points"]]
        fs = O.fit_sphere(tv, tn, base)
        assert (fs is not None) == case["sphere"]["accept"], case["cite"]
        if "center" in case["sphere"]:
            np.testing.assert_allclose(fs.a, case["sphere"]["center"], atol=1e-15)
            assert fs.s == pytest.approx(case["sphere"]["radius"], abs=1e-15)
            assert fs.outwards is case["sphere"]["outwards"]
        if "sphere_eps_0.01" in case:
            assert O.fit_sphere(tv, tn, O.ransacparameters(base, sphere={"eps": 0.01})) is None
        if "sphere_eps_10_alpha_pi_2" in case:
            assert O.fit_sphere(tv, tn, O.ransacparameters(base, sphere={"eps": 10, "alpha": math.pi / 2})) is None
        assert (O.fit_plane(tv, tn, prp) is not None) == case["plane_accept"], case["cite"]


def test_reference_confidence_and_defaults(known):
    c = known["confidenceintervals"]
    ci = O.ConfidenceInterval(c["ctor"]["min"], c["ctor"]["max"])
    assert ci.E == c["ctor"]["E"]
    with pytest.raises(ValueError):
        O.ConfidenceInterval(c["ctor"]["max"], c["ctor"]["min"])
    n = c["notsoconfident"]
    nc = O.notsoconfident(n["x"], n["y"])
    assert (nc.min, nc.max, nc.E) == (n["min"], n["max"], n["E"])
    d = known["defaults"]
    p = O.default_parameters()
    for k in ("drawN", "minsubsetN", "prob_det", "tau", "itermax", "extract_s", "terminate_s"):
        assert p["iteration"][k] == d["iteration"][k]
    assert [O.SHAPE_NAMES[t] for t in p["iteration"]["shape_types"]] == d["iteration"]["shape_types"]
    assert p["common"] == d["common"]
    for name in ("plane", "sphere", "cylinder", "cone"):
        assert p[name]["eps"] == d[name]["eps"]
        assert p[name]["alpha"] == math.radians(d[name]["alpha_deg"])
    assert p["sphere"]["sphere_par"] == d["sphere"]["sphere_par"]
    assert p["cone"]["minconeopang"] == math.radians(d["cone"]["minconeopang_deg"])


def _shapes(small):
    return [O.shape_from_params7(int(t), bool(o), list(p)) for t, o, p in zip(small["cand_type"], small["cand_outwards"], small["cand_p7"])]


def test_numpy_oracle_reproduces_masks(small):
    P, N = small["vertices"].astype(np.float64), small["normals"].astype(np.float64)
    want = np.unpackbits(small["masks"], axis=1, bitorder="little")[:, : len(P)].astype(bool)
    op = O.default_parameters()
    for i, sh in enumerate(_shapes(small)):
        np.testing.assert_array_equal(O.compatibles(sh, P, N, op), want[i])


def test_c_oracle_reproduces_masks_and_fits(small):
    from oracle import c_oracle as CO

    if not CO.available():
        pytest.skip("liboracle.so not built")
    P, N = small["vertices"].astype(np.float64), small["normals"].astype(np.float64)
    want = np.unpackbits(small["masks"], axis=1, bitorder="little")[:, : len(P)].astype(bool)
    op = O.default_parameters()
    counts, _, masks = CO.score_counts(_shapes(small), P, N, op, want_masks=True)
    np.testing.assert_array_equal(masks, want)
    np.testing.assert_array_equal(counts, want.sum(1))
    sets = small["fit_sets"]
    shapes, sset = CO.fit_points(P[sets], N[sets], op)  # [(type, outwards, p[7..])]
    assert [t for t, _, _ in shapes] == small["fit_kind"].tolist()
    assert list(sset) == small["fit_set"].tolist()
    for (t, o, p), wo, wp in zip(shapes, small["fit_outwards"], small["fit_p7"]):
        if t != 0:
            assert int(o) == int(wo)
        np.testing.assert_allclose(np.asarray(p)[:7], wp, rtol=1e-9, atol=1e-9)


def test_numpy_oracle_reproduces_loop(small):
    P32, N32 = small["vertices"], small["normals"]
    tau, msn, itmax, seed = (int(x) for x in small["run_iteration"])
    op = O.ransacparameters(O.default_parameters(), iteration={"tau": tau, "minsubsetN": msn, "itermax": itmax})
    ex = O.ransac(O.Cloud(P32, N32, [small["subset0"].copy(), small["subset1"].copy()]), op, True, seed=seed)
    assert [e.shape.kind for e in ex] == small["run_kind"].tolist()
    assert [len(e.inpoints) for e in ex] == small["run_len"].tolist()
    np.testing.assert_array_equal(np.concatenate([e.inpoints for e in ex]), small["run_inpoints"])
