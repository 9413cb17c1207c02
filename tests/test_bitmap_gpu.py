"""rsc_bitmap_filter (extension, SURVEY 8(f)-4 second half) against oracle/ransac_oracle.py::bitmap_filter:
the same points kept, the same grid and component statistics -- on every shape type, with several patches,
across the azimuth seam, with 4- and 8-connectivity, and on inlier lists of real refits."""
import math

import numpy as np
import pytest

from oracle import ransac_oracle as O
from tests.helpers import to_oracle_shape

pytestmark = pytest.mark.gpu


def _f32(a):
    return np.asarray(a, np.float64).astype(np.float32)


def _check(R, shape, P, beta, eight):
    P = _f32(P)
    N = np.zeros_like(P)
    N[:, 2] = 1
    pc = R.RANSACCloud(P, N, [np.zeros(0, np.int64)])
    idx = np.arange(len(P), dtype=np.int64)
    got, info = R.bitmap_filter(shape, pc, idx, beta, eight)
    want, (nu, nv, ncomp, cells) = O.bitmap_filter(to_oracle_shape(shape), P.astype(np.float64), beta, eight)
    assert (info["nu"], info["nv"], info["components"], info["largest_cells"]) == (nu, nv, ncomp, cells)
    np.testing.assert_array_equal(got, want)
    pc.close()
    return got, ncomp


@pytest.mark.parametrize("eight", [False, True], ids=["4conn", "8conn"])
def test_patches_on_every_shape_type(eight):
    import ransac_jl_b200 as R

    rng = np.random.default_rng(17)
    # plane through (1,2,3) with a tilted normal: a large and two small patches
    nrm = np.array([0.3, -0.5, 0.8]); nrm /= np.linalg.norm(nrm)
    plane = R.FittedPlane(np.array([1.0, 2.0, 3.0]), nrm)
    b = np.cross(nrm, [1, 0, 0]); b /= np.linalg.norm(b); c = np.cross(nrm, b)
    def patch(u0, v0, w, n):
        uv = rng.uniform(0, w, (n, 2)) + [u0, v0]
        return plane.point + uv[:, :1] * b + uv[:, 1:] * c
    kept, ncomp = _check(R, plane, np.vstack([patch(0, 0, 12, 6000), patch(30, 5, 3, 500), patch(-20, -20, 2, 300)]), 0.6, eight)
    assert ncomp == 3 and len(kept) == 6000
    # sphere: a cap around the seam (phi = pi) and a small cap elsewhere
    sph = R.FittedSphere(np.array([5.0, -1.0, 2.0]), 4.0, True)
    def cap(phi0, th0, w, n):
        phi, th = rng.uniform(phi0 - w, phi0 + w, n), rng.uniform(th0 - w / 2, th0 + w / 2, n)
        return sph.center + 4.0 * np.c_[np.cos(th) * np.cos(phi), np.cos(th) * np.sin(phi), np.sin(th)]
    kept, ncomp = _check(R, sph, np.vstack([cap(math.pi, 0.2, 0.5, 5000), cap(0.3, -0.4, 0.15, 400)]), 0.25, eight)
    assert ncomp == 2 and len(kept) == 5000
    # cylinder with a slanted axis: a band across the seam + a blob
    ax = np.array([0.2, 0.9, 0.4]); ax /= np.linalg.norm(ax)
    ctr = np.array([3.0, 0.0, -2.0]); ctr = ctr - ax * float(ax @ ctr)
    cyl = R.FittedCylinder(ax, ctr, 2.5, True)
    ox = O.normalize3(O.stable_orthogonal(ax)); oy = O.normalize3(O.cross3(ax, ox))
    def band(phi_lo, phi_hi, h0, h1, n):
        phi, h = rng.uniform(phi_lo, phi_hi, n), rng.uniform(h0, h1, n)
        return ctr + 2.5 * (np.cos(phi)[:, None] * ox + np.sin(phi)[:, None] * oy) + h[:, None] * ax
    pts = np.vstack([band(2.4, math.pi, 0, 4, 2500), band(-math.pi, -2.4, 0, 4, 2500), band(-0.2, 0.2, 9, 9.6, 300)])
    kept, ncomp = _check(R, cyl, pts, 0.3, eight)
    assert ncomp == 2 and len(kept) == 5000
    # cone: a ring section and a small patch nearer to the apex
    cone = R.FittedCone(np.array([0.0, 0.0, 10.0]), np.array([0.0, 0.0, -1.0]), math.radians(50), True)
    def cpatch(phi0, w, s0, s1, n):
        phi, s = rng.uniform(phi0 - w, phi0 + w, n), rng.uniform(s0, s1, n)
        half = math.radians(25)
        ox2 = O.normalize3(O.stable_orthogonal(np.array([0.0, 0.0, -1.0]))); oy2 = O.normalize3(O.cross3(np.array([0.0, 0.0, -1.0]), ox2))
        rad = np.cos(phi)[:, None] * ox2 + np.sin(phi)[:, None] * oy2
        return cone.apex + s[:, None] * (math.cos(half) * np.array([0.0, 0.0, -1.0]) + math.sin(half) * rad)
    kept, ncomp = _check(R, cone, np.vstack([cpatch(3.0, 0.8, 6, 12, 5000), cpatch(0.0, 0.1, 2, 2.5, 200)]), 0.3, eight)
    assert ncomp == 2 and len(kept) == 5000


def test_diagonal_chain_needs_eight_connectivity():
    import ransac_jl_b200 as R

    plane = R.FittedPlane(np.zeros(3), np.array([0.0, 0.0, 1.0]))
    ox = O.normalize3(O.stable_orthogonal(np.array([0.0, 0.0, 1.0]))); oy = O.normalize3(O.cross3(np.array([0.0, 0.0, 1.0]), ox))
    k = np.arange(40)
    chain = (k[:, None] + 0.5) * ox + (k[:, None] + 0.5) * oy     # one point per diagonal cell
    block = np.array([[60 + i + 0.5, 5 + j + 0.5] for i in range(4) for j in range(4)])
    block = block[:, :1] * ox + block[:, 1:] * oy
    corners = np.array([0.0 * ox, 64.0 * ox + 40.0 * oy])         # pin the bounding box: cells of size 1
    P = np.vstack([corners, chain, block])
    k4, _ = _check(R, plane, P, 1.0, False)
    k8, _ = _check(R, plane, P, 1.0, True)
    assert len(k4) == 16 and len(k8) >= 40


def test_on_refit_inliers_of_a_noisy_scene():
    """the use it is meant for: filter the inlier list of a refit"""
    import ransac_jl_b200 as R
    from ransac_jl_b200 import scenes
    from tests.helpers import oracle_params

    sc = scenes.scene_mixed(29, 200_000, noise_frac=0.003, jitter_deg=1.0, outlier_frac=0.3, counts=(2, 1, 1, 1))
    pc = R.RANSACCloud(sc.vertices, sc.normals, [np.zeros(0, np.int64)])
    params = R.ransacparameters()
    for prim in sc.primitives:
        ex = R.refit(prim.shape, pc, params)
        assert len(ex.inpoints) > 1000
        for eight in (False, True):
            got, info = R.bitmap_filter(prim.shape, pc, ex.inpoints, 1.0, eight)
            P = sc.vertices[ex.inpoints].astype(np.float64)
            want, stats = O.bitmap_filter(to_oracle_shape(prim.shape), P, 1.0, eight)
            np.testing.assert_array_equal(got, ex.inpoints[want])
            assert (info["nu"], info["nv"], info["components"], info["largest_cells"]) == stats
            assert len(got) > 0.8 * len(ex.inpoints)  # the primitive itself is one piece; outliers near its surface are not
    pc.close()


def test_loop_with_the_bitmap_switch_matches_the_oracle_loop():
    """RSC_EXTRACT_BITMAP inside rsc_ransac_run on a scene with two separate coplanar patches: without the filter
    they are extracted as one plane, with it the large patch, the sphere and then the small patch -- and the
    device loop equals the oracle's loop with the same filter (shapes, inlier lists, isenabled)"""
    import ransac_jl_b200 as R
    from tests.helpers import oracle_params

    rng = np.random.default_rng(7)
    A = np.c_[rng.uniform(0, 20, (12000, 2)), rng.normal(0, 0.02, 12000)]
    B = np.c_[rng.uniform(60, 70, (3000, 2)), rng.normal(0, 0.02, 3000)]
    d = rng.normal(size=(8000, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    S = np.array([30, 30, 15.0]) + 6 * d
    out = rng.uniform(-10, 80, (5000, 3))
    no = rng.normal(size=(5000, 3))
    no /= np.linalg.norm(no, axis=1, keepdims=True)
    V = np.vstack([A, B, S, out]).astype(np.float32)
    N = np.vstack([np.tile([0, 0, 1.0], (15000, 1)), d, no]).astype(np.float32)
    perm = rng.permutation(len(V))
    V, N = V[perm], N[perm]
    pc = R.RANSACCloud(V, N, 4)
    params = R.ransacparameters(iteration={"tau": 300, "minsubsetN": 128, "itermax": 80})
    plain, _ = R.ransac(pc, params, True, seed=5)
    assert [len(e.inpoints) for e in plain[:2]] == [15000, 8000]
    got, _ = R.ransac(pc, params, True, seed=5, bitmap=(1.5, False))
    oc = O.Cloud(V, N, [s.copy() for s in pc.subsets])
    want = O.ransac(oc, oracle_params(params), True, seed=5, bitmap=(1.5, False))
    assert [len(e.inpoints) for e in want] == [12000, 8000, 3000]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g.shape.to_cand().type == w.shape.kind
        np.testing.assert_array_equal(g.inpoints, w.inpoints)
    np.testing.assert_array_equal(pc.isenabled, oc.isenabled)
