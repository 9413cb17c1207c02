"""JSON export and YAML config import against the reference's own tests (CPU only).

test/json.jl:1-107 and test/yaml.jl:62-89 restated with the same inputs and expectations; the YAML
fixtures are tests/golden/ref_yaml_t{1,2}.yml (= test/yaml/t1.yml, t2.yml)."""
import io
import json
import math
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _R():
    import ransac_jl_b200 as R

    return R


def _unit(v):
    v = np.asarray(v, float)
    return v / np.linalg.norm(v)


def _shapes(R):
    p1 = np.array([15.6, 0, -13.7])
    n1 = _unit([34, 45, 7])
    a1 = np.array([-17.1, 8, 2.42])
    ap1 = np.zeros(3)
    ax1 = _unit([-1.5, 7, 2])
    s_plane = R.FittedPlane(p1, n1)
    s_sphere = R.FittedSphere(p1, 13.23444, True)
    s_cylinder = R.FittedCylinder(a1, p1, 0.13, False)
    s_cone = R.FittedCone(ap1, ax1, 0.785, True)
    d_plane = {"type": "plane", "point": list(p1), "normal": list(n1)}
    d_sphere = {"type": "sphere", "radius": 13.23444, "center": list(p1), "outwards": True}
    d_cylinder = {"type": "cylinder", "axis": list(a1), "center": list(p1), "radius": 0.13, "outwards": False}
    d_cone = {"type": "cone", "apex": list(ap1), "axis": list(ax1), "opang": 0.785, "outwards": True}
    return (s_plane, s_sphere, s_cylinder, s_cone), (d_plane, d_sphere, d_cylinder, d_cone)


def test_todict_matches_reference_expectations():
    # test/json.jl:5-57
    R = _R()
    (s_plane, s_sphere, s_cylinder, s_cone), (d_plane, d_sphere, d_cylinder, d_cone) = _shapes(R)
    assert R.toDict(s_plane) == d_plane
    assert R.toDict(s_sphere) == d_sphere
    assert R.toDict(s_cylinder) == d_cylinder
    assert R.toDict(s_cone) == d_cone
    sc1 = R.ExtractedShape(s_plane, np.array([1]))
    ss1 = R.ExtractedShape(s_cone, np.array([1, 2, 3]))
    assert R.toDict(sc1) == d_plane and R.toDict(ss1) == d_cone
    assert R.toDict([s_plane, s_sphere, s_cylinder, s_cone]) == {"primitives": [d_plane, d_sphere, d_cylinder, d_cone]}
    assert R.toDict([sc1, ss1]) == {"primitives": [d_plane, d_cone]}


def test_exportjson_round_trips(tmp_path):
    # test/json.jl:59-106: export (with / without indentation, to buffers and files), parse, compare
    R = _R()
    (s_plane, s_sphere, s_cylinder, s_cone), (d_plane, d_sphere, d_cylinder, d_cone) = _shapes(R)
    b = io.StringIO()
    R.exportJSON(b, s_cone, 2)
    assert json.loads(b.getvalue()) == d_cone and "\n  " in b.getvalue()
    b = io.StringIO()
    R.exportJSON(b, [s_plane, s_cone], 1)
    assert json.loads(b.getvalue()) == {"primitives": [d_plane, d_cone]}
    b = io.StringIO()
    R.exportJSON(b, s_cylinder)
    assert json.loads(b.getvalue()) == d_cylinder and "\n" not in b.getvalue() and ": " not in b.getvalue()
    b = io.StringIO()
    R.exportJSON(b, [s_plane, s_sphere, s_cone, s_plane])
    assert json.loads(b.getvalue()) == {"primitives": [d_plane, d_sphere, d_cone, d_plane]}
    f1 = tmp_path / "a.json"
    with open(f1, "w") as f:
        R.exportJSON(f, [s_plane, s_sphere, s_cone, s_plane], 4)
    assert json.loads(f1.read_text()) == {"primitives": [d_plane, d_sphere, d_cone, d_plane]}
    f2 = tmp_path / "b.json"
    with open(f2, "w") as f:
        R.exportJSON(f, [s_plane, s_sphere, s_cone, s_plane, s_cylinder])
    assert json.loads(f2.read_text()) == {"primitives": [d_plane, d_sphere, d_cone, d_plane, d_cylinder]}


def test_readconfig_t1():
    # test/yaml.jl:62-75
    R = _R()
    conf = R.readconfig(os.path.join(GOLD, "ref_yaml_t1.yml"))
    p = R.ransacparameters()
    p = R.ransacparameters(p, sphere={"eps": 0.2, "alpha": 0.05, "sphere_par": 0.01}, plane={"eps": 0.1, "alpha": 0.01})
    p = R.ransacparameters(p, cylinder={"alpha": 0.0872}, cone={"eps": 1, "alpha": 3.14, "minconeopang": 1.0})
    p = R.ransacparameters(p, iteration={"drawN": 9, "minsubsetN": 2, "prob_det": 0.999, "tau": 10000, "itermax": 100000,
                                         "shape_types": [R.FittedPlane, R.FittedSphere]})
    p = R.ransacparameters(p, common={"parallelthrdeg": 0.5, "collin_threshold": 0.3})
    assert conf == p
    assert conf["cylinder"]["eps"] == 0.3 and conf["iteration"]["extract_s"] == "nofminset"  # untouched defaults survive


def test_readconfig_t2():
    # test/yaml.jl:77-89
    R = _R()
    conf = R.readconfig(os.path.join(GOLD, "ref_yaml_t2.yml"))
    p = R.ransacparameters()
    p = R.ransacparameters(p, plane={"eps": 0.35, "alpha": 1.0872}, sphere={"sphere_par": 0.025})
    p = R.ransacparameters(p, iteration={"itermax": 100})
    p = R.ransacparameters(p, common={"collin_threshold": 0.22, "parallelthrdeg": 1.2})
    assert conf == p
    assert conf["sphere"]["alpha"] == math.radians(5)


def test_readconfig_feeds_the_c_abi():
    """the parameters read from a config flatten into rsc_params like any other"""
    R = _R()
    cp = R.to_c(R.readconfig(os.path.join(GOLD, "ref_yaml_t1.yml"), R.ransacparameters(iteration={"drawN": 3})))
    assert (cp.drawN, cp.minsubsetN, cp.tau, cp.itermax, cp.n_shape_types) == (9, 2, 10000, 100000, 2)
    assert [cp.shape_types[i] for i in range(2)] == [0, 1]
    assert abs(cp.eps[1] - 0.2) < 1e-15 and abs(cp.alpha[3] - 3.14) < 1e-15 and cp.parallelthrdeg == 0.5
