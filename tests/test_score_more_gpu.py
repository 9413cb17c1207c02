"""More GPU parity cases for K2/K1 against the C oracle: the packed K=2 path, non-unit normals and
axes, clouds far from the origin, the float64 upload, the Q4 switch, queue overflow, fit variants."""
import math
import os

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import ransac_oracle as O
from tests.helpers import oracle_params

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def R():
    import ransac_jl_b200 as R

    return R


def _oracle_counts(cands, P, N, op, enabled=None, sphere_ignores=True):
    prm_flag = sphere_ignores
    lib = CO._load()
    # c_oracle.score_counts always applies Q4; emulate the switch by passing spheres separately
    counts, _, masks = CO.score_counts(cands, P, N, op, enabled=enabled, want_masks=True)
    if not prm_flag and enabled is not None:
        for i, c in enumerate(cands):
            if c.kind == 1:
                masks[i] &= enabled
                counts[i] = masks[i].sum()
    return counts, masks


def test_k2_path_1000_candidates(R):
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(101, 60_000)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 1)
    params = R.ransacparameters()
    cands = scenes.perturbed_candidates(sc, 250, seed=9)  # 1000 candidates -> K = 2 (one packed pair per thread)
    en = np.random.default_rng(2).random(pc.size) > 0.25
    pc.isenabled = en
    counts, masks = R.score_counts(pc, cands, -1, params, want_masks=True)
    P, N = sc.vertices.astype(np.float64), sc.normals.astype(np.float64)
    want, wmask = _oracle_counts(cands, P, N, oracle_params(params), enabled=en)
    np.testing.assert_array_equal(counts, want)
    for i in range(0, len(cands), 37):
        np.testing.assert_array_equal(R.unpack_mask(masks[i], pc.size), wmask[i])


def test_non_unit_normals_and_axes(R):
    """the reference never normalises in the scorers (Q15/Q18): scaled point normals, scaled plane
    normals and cylinder/cone axes must give the reference's (odd) answers, bit for bit"""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(102, 30_000)
    rng = np.random.default_rng(4)
    nrm = (sc.normals.astype(np.float64) * rng.uniform(0.5, 1.5, (len(sc.normals), 1))).astype(np.float32)
    pc = R.RANSACCloud(sc.vertices, nrm, 1)
    params = R.ransacparameters()
    base = [p.shape for p in sc.primitives] + scenes.perturbed_candidates(sc, 6, seed=3)
    cands = []
    for sh in base:
        k = rng.uniform(0.7, 1.3)
        if isinstance(sh, R.FittedPlane):
            cands.append(R.FittedPlane(sh.point, sh.normal * k))
        elif isinstance(sh, R.FittedCylinder):
            cands.append(R.FittedCylinder(sh.axis * k, sh.center, sh.radius, sh.outwards))
        elif isinstance(sh, R.FittedCone):
            cands.append(R.FittedCone(sh.apex, sh.axis * k, sh.opang, sh.outwards))
        else:
            cands.append(sh)
    counts, masks = R.score_counts(pc, cands, -1, params, want_masks=True)
    want, wmask = _oracle_counts(cands, sc.vertices.astype(np.float64), nrm.astype(np.float64), oracle_params(params))
    np.testing.assert_array_equal(counts, want)
    for i in range(len(cands)):
        np.testing.assert_array_equal(R.unpack_mask(masks[i], pc.size), wmask[i])


def test_cloud_far_from_origin(R):
    """coordinates ~1e4: the FP32 band widens, more pairs go to FP64, answers stay exact"""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(103, 20_000)
    shift = np.array([8000.0, -12000.0, 5000.0])
    P32 = (sc.vertices.astype(np.float64) + shift).astype(np.float32)
    pc = R.RANSACCloud(P32, sc.normals, 1)
    params = R.ransacparameters()
    cands = []
    for sh in [p.shape for p in sc.primitives]:
        c = sh.to_cand()
        p = list(c.p)
        if c.type == 2:  # cylinder centre lies on the plane through the origin: re-project after the shift
            a = np.array(p[0:3]); q = np.array(p[3:6]) + shift; p[3:6] = list(q - a * float(a @ q))
        else:
            p[0:3] = list(np.array(p[0:3]) + shift)
        c.p[:] = p
        cands.append(R.from_cand(c))
    before = pc.ctx.stats().exact_pairs
    counts, masks = R.score_counts(pc, cands, -1, params, want_masks=True)
    want, wmask = _oracle_counts(cands, P32.astype(np.float64), sc.normals.astype(np.float64), oracle_params(params))
    np.testing.assert_array_equal(counts, want)
    for i in range(len(cands)):
        np.testing.assert_array_equal(R.unpack_mask(masks[i], pc.size), wmask[i])
    print("fp64 pairs:", pc.ctx.stats().exact_pairs - before, "of", len(cands) * pc.size)


def test_float64_upload_rounds_to_float32(R):
    from ransac_jl_b200 import scenes

    sc = scenes.scene_c1(seed=21)
    rng = np.random.default_rng(1)
    P64 = sc.vertices.astype(np.float64) + rng.normal(scale=1e-9, size=sc.vertices.shape)  # not float32-representable
    N64 = sc.normals.astype(np.float64)
    pc = R.RANSACCloud(P64, N64, 1)  # rsc_cloud_create_f64
    params = R.ransacparameters()
    cands = [p.shape for p in sc.primitives]
    counts, _ = R.score_counts(pc, cands, -1, params)
    want, _ = _oracle_counts(cands, P64.astype(np.float32).astype(np.float64), N64.astype(np.float32).astype(np.float64), oracle_params(params))
    np.testing.assert_array_equal(counts, want)


def test_sphere_quirk_switch(R):
    """compat_flags = 0: spheres honour `isenabled` like the other shapes (Q4 switched off)"""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_c1(seed=22)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 1)
    en = np.random.default_rng(3).random(pc.size) > 0.5
    pc.isenabled = en
    params = R.ransacparameters()
    cands = [p.shape for p in sc.primitives]
    P, N = sc.vertices.astype(np.float64), sc.normals.astype(np.float64)
    c_q4, _ = R.score_counts(pc, cands, -1, params)
    c_fix, _ = R.score_counts(pc, cands, -1, params, compat_flags=0)
    w_q4, _ = _oracle_counts(cands, P, N, oracle_params(params), enabled=en, sphere_ignores=True)
    w_fix, _ = _oracle_counts(cands, P, N, oracle_params(params), enabled=en, sphere_ignores=False)
    np.testing.assert_array_equal(c_q4, w_q4)
    np.testing.assert_array_equal(c_fix, w_fix)
    assert c_q4[1] > c_fix[1]  # the sphere loses its disabled inliers only when the quirk is off


def test_guard_queue_overflow_grows_and_retries(R):
    """a context that starts with a 64-entry guard-band queue still returns exact counts"""
    from ransac_jl_b200 import scenes

    os.environ["RSC_WL_CAP"] = "64"
    try:
        ctx = R.Context(0)
    finally:
        del os.environ["RSC_WL_CAP"]
    sc = scenes.scene_mixed(104, 80_000)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 1, ctx=ctx)
    params = R.ransacparameters()
    cands = [p.shape for p in sc.primitives] + scenes.perturbed_candidates(sc, 20, seed=5)
    counts, _ = R.score_counts(pc, cands, -1, params)
    want, _ = _oracle_counts(cands, sc.vertices.astype(np.float64), sc.normals.astype(np.float64), oracle_params(params))
    np.testing.assert_array_equal(counts, want)
    assert ctx.stats().exact_pairs > 64


def test_fit_four_point_sets_and_type_subsets(R):
    """drawN = 4 (the 4th point only validates, like test/dummyspheretest.jl) and other shape_types"""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(105, 40_000, noise_frac=0.002)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 1)
    rng = np.random.default_rng(6)
    lab = sc.labels
    pools = [np.flatnonzero(lab == l) for l in range(lab.max() + 1)]
    idx = np.stack([rng.choice(pools[rng.integers(len(pools))], 4, replace=False) for _ in range(3000)])
    P, N = sc.vertices.astype(np.float64), sc.normals.astype(np.float64)
    for types in ([R.FittedSphere], [R.FittedCone, R.FittedPlane], [R.FittedCylinder, R.FittedSphere, R.FittedPlane]):
        params = R.ransacparameters(types, iteration={"drawN": 4})
        shapes, sets = R.fit_batch(pc, idx, params)
        want, want_set = CO.fit_points(P[idx], N[idx], oracle_params(params))
        assert len(shapes) == len(want) and len(want) > 50, (types, len(shapes), len(want))
        np.testing.assert_array_equal(sets, want_set)
        for sh, (t, outw, p) in zip(shapes, want):
            c = sh.to_cand()
            assert c.type == t and bool(c.outwards) == outw
            assert np.abs(np.array(c.p[:]) - p).max() <= 1e-5 * max(1.0, np.abs(p).max())


def test_wide_and_flat_cones(R):
    """the tiled kernel evaluates cones divided by cos(opang/2); cones wider than 172.8 deg (and
    opang >= 180 deg, where that cosine is <= 0) are routed to the FP64 path wholesale.  Counts,
    masks and refit lists must still be the reference's."""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(105, 20_000)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 1)
    params = R.ransacparameters()
    rng = np.random.default_rng(8)
    P, N = sc.vertices.astype(np.float64), sc.normals.astype(np.float64)
    cands = []
    for deg in (120.0, 160.0, 170.0, 172.0, 173.5, 178.0, 179.99, 180.0, 185.0, 270.0):
        for outw in (True, False):
            a = rng.normal(size=3)
            a /= np.linalg.norm(a)
            apex = P[rng.integers(len(P))] - a * rng.uniform(0.0, 3.0)
            cands.append(R.FittedCone(apex, a, math.radians(deg), outw))
    counts, masks = R.score_counts(pc, cands, -1, params, want_masks=True)
    want, wmask = _oracle_counts(cands, P, N, oracle_params(params))
    np.testing.assert_array_equal(counts, want)
    for i in range(len(cands)):
        np.testing.assert_array_equal(R.unpack_mask(masks[i], pc.size), wmask[i])
    assert want.sum() > 0
    for i in (0, 9, 13):
        ex = R.refit(cands[i], pc, params)
        np.testing.assert_array_equal(ex.inpoints, np.flatnonzero(wmask[i]))


def test_refit_guard_queue_overflow(R, monkeypatch):
    """refit queues the points inside the FP32 guard band for FP64; with a tiny queue the fix-up
    kernel must fall back to re-scanning the range and still return the reference's list"""
    from ransac_jl_b200 import scenes

    monkeypatch.setenv("RSC_EXQ_CAP", "8")
    sc = scenes.scene_mixed(106, 50_000)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 1)
    rng = np.random.default_rng(5)
    en = rng.random(pc.size) > 0.3
    pc.isenabled = en
    params = R.ransacparameters()
    P, N = sc.vertices.astype(np.float64), sc.normals.astype(np.float64)
    a = np.array([0.0, 0.0, 1.0])
    flat = R.FittedCone(P[10] - a * 0.5, a, math.radians(176.0), True)  # infinite band: every enabled point queues
    cands = [p.shape for p in sc.primitives][:6] + [flat]
    want, wmask = _oracle_counts(cands, P, N, oracle_params(params), enabled=en)
    for i, sh in enumerate(cands):
        ex = R.refit(sh, pc, params)
        m = wmask[i] & en
        np.testing.assert_array_equal(ex.inpoints, np.flatnonzero(m))


def test_update_coordinates_drops_cells_and_rescoring_follows(R):
    """rsc_cloud_update replaces the coordinates in place: scoring must see the new points, the
    flattened octree built for the old ones must be gone (the sampler refuses until it is rebuilt)"""
    import ctypes as C

    from ransac_jl_b200 import scenes
    from ransac_jl_b200._lib import lib

    a = scenes.scene_mixed(107, 30_000)
    b = scenes.scene_mixed(108, 30_000)
    pc = R.RANSACCloud(a.vertices, a.normals, 1).build_cells(6)
    params = R.ransacparameters()
    cands = [p.shape for p in b.primitives]
    assert lib.rsc_cloud_cells_levels(pc.handle) == 6
    pc.ctx.check(lib.rsc_cloud_update(pc.handle, b.vertices.ctypes.data, b.normals.ctypes.data, pc.size))
    counts, _ = R.score_counts(pc, cands, -1, params)
    want, _ = _oracle_counts(cands, b.vertices.astype(np.float64), b.normals.astype(np.float64), oracle_params(params))
    np.testing.assert_array_equal(counts, want)
    assert lib.rsc_cloud_cells_levels(pc.handle) == 0
    with pytest.raises(R.RscError):
        R.sample_fit_cells(pc, params, 1, 0, 16, np.full(6, 1 / 6))


def test_many_candidates_multi_column_masks(R):
    """20 000 candidates (several CTA columns per type, K = 4 tiling) with masks on a small cloud"""
    from ransac_jl_b200 import scenes

    sc = scenes.scene_mixed(109, 4_000)
    pc = R.RANSACCloud(sc.vertices, sc.normals, 1)
    params = R.ransacparameters()
    cands = scenes.perturbed_candidates(sc, 5000, seed=13)
    counts, masks = R.score_counts(pc, cands, -1, params, want_masks=True)
    want, wmask = _oracle_counts(cands, sc.vertices.astype(np.float64), sc.normals.astype(np.float64), oracle_params(params))
    np.testing.assert_array_equal(counts, want)
    for i in range(0, len(cands), 211):
        np.testing.assert_array_equal(R.unpack_mask(masks[i], pc.size), wmask[i])


def test_score_dev_masks_device_pointers(R):
    """rsc_score_dev_masks: candidates, counts and candidate-major masks all in device memory"""
    import ctypes as C

    import torch

    from ransac_jl_b200 import scenes
    from ransac_jl_b200._lib import lib

    sc = scenes.scene_mixed(110, 20_001)  # ragged: not a multiple of 32
    pc = R.RANSACCloud(sc.vertices, sc.normals, 2)
    en = np.random.default_rng(1).random(pc.size) > 0.3
    pc.isenabled = en
    params = R.ransacparameters()
    cands = [p.shape for p in sc.primitives] + scenes.perturbed_candidates(sc, 40, seed=3)
    cp = R.to_c(params)
    dev = torch.device("cuda", 0)
    d_cands = torch.frombuffer(bytearray(bytes(R.pack_cands(cands))), dtype=torch.uint8).to(dev)
    for subset in (-1, 0):
        m = pc.size if subset < 0 else len(pc.subsets[0])
        words = (m + 31) // 32
        d_counts = torch.zeros(len(cands), dtype=torch.int32, device=dev)
        d_masks = torch.zeros((len(cands), words), dtype=torch.int32, device=dev)
        pc.ctx.check(lib.rsc_score_dev_masks(pc.handle, C.byref(cp), d_cands.data_ptr(), len(cands), subset, d_counts.data_ptr(),
                                             d_masks.data_ptr(), None))
        torch.cuda.synchronize()
        counts, masks = R.score_counts(pc, cands, subset, params, want_masks=True)  # host-buffer path
        np.testing.assert_array_equal(d_counts.cpu().numpy(), counts)
        np.testing.assert_array_equal(d_masks.cpu().numpy().view(np.uint32), np.asarray(masks).reshape(len(cands), words))


def test_adversarial_candidates_match_oracle(R):
    """tests.helpers.adversarial_case: every shape type with wide/flat cones, non-unit axes/normals, tiny and
    huge radii, points on axes and at centres, zero normals, NaN/Inf/zero-axis candidates -- counts, masks
    and refit lists must be the oracle's, bit for bit"""
    from ransac_jl_b200 import _lib
    from ransac_jl_b200.shapes import from_cand
    from tests.helpers import adversarial_case

    oshapes, P, N = adversarial_case()
    cands = []
    for sh in oshapes:
        c = _lib.rsc_cand(type=int(sh.kind), outwards=int(bool(sh.outwards)))
        for i, v in enumerate(sh.params7()):
            c.p[i] = float(v)
        cands.append(from_cand(c))
    pc = R.RANSACCloud(P.astype(np.float32), N.astype(np.float32), 1)
    params = R.ransacparameters()
    op = oracle_params(params)
    counts, masks = R.score_counts(pc, cands, -1, params, want_masks=True)
    want, _, wmask = CO.score_counts(oshapes, P, N, op, want_masks=True)
    np.testing.assert_array_equal(counts, want)
    for i in range(len(cands)):
        np.testing.assert_array_equal(R.unpack_mask(masks[i], pc.size), wmask[i], err_msg=f"candidate {i} kind {oshapes[i].kind}")
    for i in range(0, len(cands), 7):
        np.testing.assert_array_equal(R.refit(cands[i], pc, params).inpoints, np.flatnonzero(wmask[i]))


def test_float64_cloud_decides_on_the_float32_roundings():
    """rsc_cloud_create_f64 rounds the coordinates to float32 on upload (documented in include/rsc.h, README,
    INTEGRATION.md): masks and counts are the float64 oracle's on the ROUNDED coordinates, bit for bit -- and
    differ from the oracle on the unrounded ones only for points within float32 resolution of a threshold."""
    import ransac_jl_b200 as R
    from ransac_jl_b200 import scenes
    from tests.helpers import oracle_mask, oracle_params

    sc = scenes.scene_mixed(61, 60_000, noise_frac=0.004, jitter_deg=2.0, outlier_frac=0.2)
    rng = np.random.default_rng(5)
    V64 = sc.vertices.astype(np.float64) + rng.uniform(-1e-6, 1e-6, sc.vertices.shape)  # not float32-representable
    N64 = sc.normals.astype(np.float64) + rng.uniform(-1e-8, 1e-8, sc.normals.shape)
    pc = R.RANSACCloud(V64, N64, [np.arange(len(V64), dtype=np.int64)])
    assert pc.vertices.dtype == np.float64
    params = R.ransacparameters()
    cands = [p.shape for p in sc.primitives] + scenes.perturbed_candidates(sc, 6, seed=2)
    counts, masks = R.score_counts(pc, cands, 0, params, want_masks=True)
    op = oracle_params(params)
    Vr, Nr = V64.astype(np.float32).astype(np.float64), N64.astype(np.float32).astype(np.float64)
    differ = 0
    for i, sh in enumerate(cands):
        got = R.unpack_mask(masks[i], len(V64))
        np.testing.assert_array_equal(got, oracle_mask(sh, Vr, Nr, op))
        differ += int((got != oracle_mask(sh, V64, N64, op)).sum())
    assert differ <= 20  # a handful of threshold cases at most
    pc.close()
