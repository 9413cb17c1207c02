"""The parameter-space bitmap functions of the reference (dead code there: src/parameterspacebitmap.jl),
restated in the oracle and pinned to the reference's OWN known-answer tests -- test/parameterspacebitmap.jl,
which the reference ships but has commented out of runtests.jl (its `label_components` dependency,
ImageMorphology, is not in Project.toml)."""
import numpy as np
import pytest

from oracle import ransac_oracle as O


def _indmap(n):
    return [[[] for _ in range(n)] for _ in range(n)]


def test_largestconncomp_on_dense_bitmap():
    """test/parameterspacebitmap.jl:1-22 (1-based ranges -> 0-based slices)"""
    pic = np.zeros((150, 150), bool)
    pic[54:75, 54:75] = True
    pic[99:125, 99:125] = True
    pic[129:140, 129:140] = True
    ind = _indmap(150)
    for i in range(54, 75):
        for j in range(54, 75):
            ind[i][j] += [1, 2, 3]
    for i in range(99, 125):
        for j in range(99, 125):
            ind[i][j] += [-99, -98]
    for i in range(129, 140):
        for j in range(129, 140):
            ind[i][j] += [0]
    t1 = O.largestconncomp(pic, ind, range(1, 3))
    assert t1 == [-99, -98] * (26 * 26)
    assert t1 == O.largestconncomp(pic, ind)
    assert t1 == O.largestconncomp(pic, ind, np.ones((3, 3), bool))


def test_eight_connectivity():
    """test/parameterspacebitmap.jl:24-60: the largest patch is only 8-connected"""
    pic = np.zeros((150, 150), bool)
    pic[54:73, 54:73] = True
    pic[99:126:2, 99:126:2] = True
    pic[100:127:2, 100:127:2] = True
    pic[129:140, 129:140] = True
    ind = _indmap(150)
    for i in range(54, 73):
        for j in range(54, 73):
            ind[i][j] += [1, 2, 3]
    for i in list(range(99, 126, 2)):
        for j in range(99, 126, 2):
            ind[i][j] += [-99, -98]
    for i in range(100, 127, 2):
        for j in range(100, 127, 2):
            ind[i][j] += [-99, -98]
    for i in range(129, 140):
        for j in range(129, 140):
            ind[i][j] += [0]
    t24 = O.largestconncomp(pic, ind, range(1, 3))
    assert t24 == [1, 2, 3] * (19 * 19)
    assert O.largestconncomp(pic, ind) == t24 and O.largestconncomp(pic, ind, "default") == t24
    t28 = O.largestconncomp(pic, ind, np.ones((3, 3), bool))
    assert t28 == [-99, -98] * (14 * 14 * 2)
    assert t28 == O.largestconncomp(pic, ind, "eight")
    with pytest.raises(ValueError):
        O.largestconncomp(pic, ind, "nine")


def test_bitmapparameters_literal_quirks():
    """parameterspacebitmap.jl:12-46: box widened by 0.1, first id wins, border places dropped"""
    P = np.array([[0.0, 0.0], [0.05, 0.05], [1.0, 1.0], [2.0, 3.0], [9.0, 9.0]])
    bm, idm, (bx, by) = O.bitmapparameters(P, [True, True, True, False, True], 1.0)
    assert bm.shape == (9, 9) and bx == pytest.approx(9.2 / 9)
    assert bm.sum() == 2 and idm[0, 0] == 1  # points 1 and 2 share a cell: the first id stays
    assert idm[1, 1] == 3
    assert not bm[8, 8]  # the point at the maximum lands on place xs: dropped ("boundserror")


def test_arbitrary_orthogonal_and_project2plane():
    for v in ([0, 0, 1.0], [1, 2, 3.0], [-3, 0.5, 0.2], [0.3, 0.3, 0.3]):
        o = O.arbitrary_orthogonal(v)
        assert abs(float(np.dot(o, v))) < 1e-12 and np.linalg.norm(o) > 0.1
    sh = O.shape_from_params7(0, True, [1, 2, 3, 0, 0, 2.0, 0])
    q = O.project2plane(sh, np.array([[1, 2, 3.0], [2, 2, 3.0], [1, 2, 5.0]]))
    assert np.allclose(q[0], 0) and abs(q[2, 2] - 2.0) < 1e-12 and abs(np.hypot(q[1, 0], q[1, 1]) - 1.0) < 1e-12


def test_bitmap_filter_keeps_the_largest_patch_and_wraps_the_seam():
    rng = np.random.default_rng(3)
    # plane z = 0: a big patch and a small far one
    big = np.c_[rng.uniform(0, 10, (4000, 2)), np.zeros(4000)]
    small = np.c_[rng.uniform(30, 32, (300, 2)), np.zeros(300)]
    pts = np.vstack([big, small])
    plane = O.shape_from_params7(0, True, [0, 0, 0, 0, 0, 1.0, 0])
    keep, (nu, nv, ncomp, cells) = O.bitmap_filter(plane, pts, 0.5)
    assert ncomp == 2 and set(keep) == set(range(4000))
    # cylinder about z, radius 2: one band of points that crosses the seam phi = +-pi, plus a small blob
    phi = np.r_[rng.uniform(2.2, np.pi, 1500), rng.uniform(-np.pi, -2.2, 1500)]
    band = np.c_[2 * np.cos(phi), 2 * np.sin(phi), rng.uniform(0, 3, 3000)]
    phi2 = rng.uniform(-0.2, 0.2, 200)
    blob = np.c_[2 * np.cos(phi2), 2 * np.sin(phi2), rng.uniform(10, 10.5, 200)]
    cyl = O.shape_from_params7(2, True, [0, 0, 1.0, 0, 0, 0, 2.0])
    keep, (nu, nv, ncomp, cells) = O.bitmap_filter(cyl, np.vstack([band, blob]), 0.25)
    assert ncomp == 2 and set(keep) == set(range(3000)), (ncomp, len(keep))
