"""CPU oracle: literal float64 restatement of RANSAC.jl's hot path (NumPy).

TEST INFRASTRUCTURE ONLY.  Nothing in the shipped package may import this module; only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` use it, and only as the checker.

What is restated (all citations relative to /root/reference, RANSAC.jl v0.6.0):

  fit(FittedPlane)        src/shapes/plane.jl:33-57
  compatiblesPlane        src/shapes/plane.jl:82-130   (project2plane :82-103)
  fit2pointsphere / fit   src/shapes/sphere.jl:29-75, :87-114
  compatiblesSphere       src/shapes/sphere.jl:144-172
  fit2pointcylinder / fit src/shapes/cylinder.jl:34-125, :135-168
  compatiblesCylinder     src/shapes/cylinder.jl:194-221
  fit3pointcone           src/shapes/cone.jl:39-61
  project2cone            src/shapes/cone.jl:68-85     (rodrigues: src/utilities.jl:19-43,61-64)
  validatecone / fit      src/shapes/cone.jl:87-115, :123-128
  compatiblesCone         src/shapes/cone.jl:132-153
  scorecandidate x4       plane.jl:61-71 sphere.jl:118-134 cylinder.jl:172-183 cone.jl:155-167
  refit x4                plane.jl:137-143 sphere.jl:179-190 cylinder.jl:228-234 cone.jl:176-182
  estimatescore           src/confidenceintervals.jl:53-74 (Int64 wrap-around reproduced)
  findhighestscore        src/fitting.jl:140-158
  prob / chooseS          src/utilities.jl:262, :297-300
  samplepointcloud4!      src/fitting.jl:383-430 (root cell only, see Q1 below)
  ransac loop             src/iterations.jl:35-162
  RANSACCloud subsets     src/octree.jl:126-138

Parity status: the reference's own tests pin only the sphere/plane accept-reject answers of
test/dummyspheretest.jl:14-48, ConfidenceInterval (test/confidenceintervals.jl) and the default
parameters (test/utilitytests.jl:41-82); those are reproduced in tests/test_oracle_golden.py.
compatibles*/scorecandidate/refit/estimatescore and the cylinder/cone fits are pinned by NO
reference test and Julia is not installed here, so for those: PARITY UNPINNED (source
restatement only, plus analytic on-surface self-checks).  Third-party arithmetic whose exact
rounding is not restated: StaticArrays dot/cross/normalize (taken as left-to-right sums and
inv(norm)*v), LinearAlgebra rank (SVD, tol = min(m,n)*eps*smax) and `\\` (LAPACK LU), cosd.

Behavioural quirks kept on purpose (SURVEY.md section 8a'):
  Q1  the level weights are NaN/zero so every minimal set is drawn from the root cell
      (octree.jl:82-84 vs :44-45) -> the sampler here draws from all enabled points.
  Q3  plane collinearity test never fires.         Q4  sphere scoring ignores `isenabled`.
  Q5  sphere degenerate branch uses p1 twice.      Q6  cone validation distance is signed.
  Q8  cylinder fit uses raw normals.               Q9  Int64 overflow in estimatescore.
  Q12 s counts attempted minimal sets.             Q17 NaN (point on axis/centre) -> incompatible.
  Q18 plane angle test uses the raw normal, the distance the re-normalised one.

Indices are 0-based here (the reference is 1-based); everything else is op-for-op.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

F = np.float64

# --------------------------------------------------------------------------------------
# small vector helpers (StaticArrays semantics: left-to-right sums, normalize = inv(norm)*v)
# --------------------------------------------------------------------------------------


def dot3(a, b):
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]


def norm3(a):
    return np.sqrt(a[..., 0] * a[..., 0] + a[..., 1] * a[..., 1] + a[..., 2] * a[..., 2])


def normalize3(a):
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / norm3(a)
        return inv[..., None] * a if np.ndim(inv) else inv * a


def cross3(a, b):
    return np.stack(
        [
            a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1],
            a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
            a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0],
        ],
        axis=-1,
    )


def _v(x):
    return np.asarray(x, dtype=F)


# --------------------------------------------------------------------------------------
# parameters (utilities.jl:332-399; defaults pinned by test/utilitytests.jl:41-82)
# --------------------------------------------------------------------------------------

PLANE, SPHERE, CYLINDER, CONE = 0, 1, 2, 3
SHAPE_NAMES = {PLANE: "plane", SPHERE: "sphere", CYLINDER: "cylinder", CONE: "cone"}
#: RANSAC.jl:94 -- DEFAULT_PARAMETERS shape order
DEFAULT_SHAPE_TYPES = (PLANE, CONE, CYLINDER, SPHERE)


def default_parameters(shape_types: Sequence[int] = DEFAULT_SHAPE_TYPES) -> dict:
    """defaultparameters (utilities.jl:391-399) as a nested dict."""
    p = {
        "iteration": {
            "drawN": 3,
            "minsubsetN": 15,
            "prob_det": 0.9,
            "shape_types": list(shape_types),
            "tau": 900,
            "itermax": 1000,
            "extract_s": "nofminset",
            "terminate_s": "nofminset",
        },
        "common": {"collin_threshold": 0.2, "parallelthrdeg": 1.0},
    }
    for s in shape_types:
        if s == PLANE:
            p["plane"] = {"eps": 0.3, "alpha": math.radians(5)}
        elif s == SPHERE:
            p["sphere"] = {"eps": 0.3, "alpha": math.radians(5), "sphere_par": 0.02}
        elif s == CYLINDER:
            p["cylinder"] = {"eps": 0.3, "alpha": math.radians(5)}
        elif s == CONE:
            p["cone"] = {"eps": 0.3, "alpha": math.radians(5), "minconeopang": math.radians(2)}
    return p


def ransacparameters(base: Optional[dict] = None, **kw) -> dict:
    """ransacparameters (utilities.jl:425-433): per-group merge of overrides."""
    if base is None:
        base = default_parameters()
    new = {k: dict(v) for k, v in base.items()}
    for k, v in kw.items():
        old = dict(base.get(k, v))
        old.update(v)
        new[k] = old
    return new


def cosd(deg: float) -> float:
    return math.cos(math.radians(deg))


# --------------------------------------------------------------------------------------
# shapes
# --------------------------------------------------------------------------------------


@dataclass
class Shape:
    """One fitted candidate.  kind in {PLANE, SPHERE, CYLINDER, CONE}.

    plane   : point, normal                      (plane.jl:8-11)
    sphere  : center, radius, outwards           (sphere.jl:9-13)
    cylinder: axis, center, radius, outwards     (cylinder.jl:11-16)
    cone    : apex, axis, opang, outwards        (cone.jl:11-19)
    """

    kind: int
    a: np.ndarray  # plane.point | sphere.center | cylinder.axis | cone.apex
    b: np.ndarray  # plane.normal | (unused)     | cylinder.center | cone.axis
    s: float = 0.0  # radius | opang
    outwards: bool = True

    def params7(self) -> np.ndarray:
        """Flat 7-double record in the C-ABI order (include/rsc.h, rsc_cand.p)."""
        if self.kind == PLANE:
            return np.array([*self.a, *self.b, 0.0])
        if self.kind == SPHERE:
            return np.array([*self.a, self.s, 0.0, 0.0, 0.0])
        return np.array([*self.a, *self.b, self.s])


def shape_from_params7(kind: int, outwards: bool, p: Sequence[float]) -> Shape:
    p = _v(p)
    if kind == PLANE:
        return Shape(PLANE, p[0:3].copy(), p[3:6].copy(), 0.0, True)
    if kind == SPHERE:
        return Shape(SPHERE, p[0:3].copy(), np.zeros(3), float(p[3]), bool(outwards))
    return Shape(kind, p[0:3].copy(), p[3:6].copy(), float(p[6]), bool(outwards))


# ---- plane ---------------------------------------------------------------------------


def fit_plane(p, n, params) -> Optional[Shape]:
    """plane.jl:33-57."""
    p, n = _v(p), _v(n)
    alpha = params["plane"]["alpha"]
    collin = params["common"]["collin_threshold"]
    lp = len(p)
    assert lp > 2 and lp == len(n)
    crossv = normalize3(cross3(p[1] - p[0], p[2] - p[0]))
    if norm3(crossv) < collin:  # Q3: norm of a normalised vector is 1 or NaN
        return None
    thr = math.cos(alpha)
    norm_ok = np.zeros(lp, bool)
    inv_ok = np.zeros(lp, bool)
    for i in range(lp):
        d = dot3(crossv, normalize3(n[i]))
        norm_ok[i] = d > thr
        inv_ok[i] = d < -thr
    if norm_ok.all():
        return Shape(PLANE, p[0].copy(), crossv)
    if inv_ok.all():
        return Shape(PLANE, p[0].copy(), -1 * crossv)
    return None


def compatibles_plane(sh: Shape, pts, nrm, params) -> np.ndarray:
    """plane.jl:114-130 with project2plane :82-103 (only the third coordinate matters)."""
    eps = params["plane"]["eps"]
    thr = math.cos(params["plane"]["alpha"])
    o_z = normalize3(sh.b)
    v = pts - sh.a
    pz = dot3(o_z, v)
    with np.errstate(invalid="ignore"):
        return (dot3(sh.b, nrm) > thr) & (np.abs(pz) < eps)


# ---- sphere --------------------------------------------------------------------------


def fit2pointsphere(v, n, params) -> Shape:
    """sphere.jl:29-75."""
    sphere_par = params["sphere"]["sphere_par"]
    par = params["common"]["parallelthrdeg"]
    n1n = normalize3(n[0])
    n2n = normalize3(n[1])
    if abs(dot3(n1n, n2n)) > cosd(par):
        c = (v[0] + v[1]) / 2
        return Shape(SPHERE, c, np.zeros(3), float(norm3(c - v[0])), False)
    g = v[1] - v[0]
    h = cross3(n2n, g)
    k = cross3(n2n, n1n)
    nk = norm3(k)
    nh = norm3(h)
    if nk < sphere_par or nh < sphere_par:
        n2 = cross3(n2n, cross3(n1n, n2n))
        n1 = cross3(n1n, cross3(n2n, n1n))
        with np.errstate(divide="ignore", invalid="ignore"):
            c1 = v[0] + dot3(v[1] - v[0], n2) / dot3(n[0], n2) * n[0]
            c2 = v[1] + dot3(v[0] - v[1], n1) / dot3(n[1], n1) * n[1]
        c = (c1 + c2) / 2
        r = (norm3(v[0] - c) + norm3(v[0] - c)) / 2  # Q5
        return Shape(SPHERE, c, np.zeros(3), float(r), False)
    if dot3(h, k) > 0:
        m = v[0] + nh / nk * n1n
    else:
        m = v[0] - nh / nk * n1n
    return Shape(SPHERE, m, np.zeros(3), float(norm3(m - v[0])), False)


def fit_sphere(p, n, params) -> Optional[Shape]:
    """sphere.jl:87-114."""
    p, n = _v(p), _v(n)
    eps = params["sphere"]["eps"]
    alpha = params["sphere"]["alpha"]
    pl = len(p)
    assert pl == len(n) and pl > 2
    sp = fit2pointsphere(p, n, params)
    thr = math.cos(alpha)
    vert_ok = np.zeros(pl, bool)
    norm_ok = np.zeros(pl, bool)
    inv_ok = np.zeros(pl, bool)
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(pl):
            vert_ok[i] = abs(norm3(p[i] - sp.a) - sp.s) < eps
            d = dot3(normalize3(p[i] - sp.a), normalize3(n[i]))
            norm_ok[i] = d > thr
            inv_ok[i] = d < -thr
    if not vert_ok.all():
        return None
    if norm_ok.all():
        sp.outwards = True
        return sp
    if inv_ok.all():
        sp.outwards = False
        return sp
    return None


def compatibles_sphere(sh: Shape, pts, nrm, params) -> np.ndarray:
    """sphere.jl:144-172."""
    eps = params["sphere"]["eps"]
    thr = math.cos(params["sphere"]["alpha"])
    o, R = sh.a, sh.s
    with np.errstate(invalid="ignore", divide="ignore"):
        if sh.outwards:
            u = normalize3(pts - o)
        else:
            u = normalize3(o - pts)
        return (dot3(u, nrm) > thr) & (np.abs(norm3(pts - o) - R) < eps)


# ---- cylinder ------------------------------------------------------------------------


def _project2plane(n, w):
    # cylinder.jl:46-59: w + n * (dot(-n, w) / dot(n, n))
    return w + n * (dot3(-n, w) / dot3(n, n))


def _projectto2d(xa, ya, za, p):
    """cylinder.jl:61-85 (Cramer's rule, first two coordinates)."""
    xx, xy, xz = xa
    yx, yy, yz = ya
    zx, zy, zz = za
    px, py, pz = p
    den = xz * yy * zx - xy * yz * zx - xz * yx * zy + xx * yz * zy + xy * yx * zz - xx * yy * zz
    num1 = -(pz * yy * zx) + py * yz * zx + pz * yx * zy - px * yz * zy - py * yx * zz + px * yy * zz
    num2 = pz * xy * zx - py * xz * zx - pz * xx * zy + px * xz * zy + py * xx * zz - px * xy * zz
    return np.array([-(num1 / den), -(num2 / den)])


def _lineintersection(a, b, c, d):
    """cylinder.jl:87-101; det of an SMatrix{2,2} is the closed form a11*a22 - a12*a21."""
    amb = a - b
    cmd = c - d
    d1 = a[0] * b[1] - a[1] * b[0]
    d2 = c[0] * d[1] - c[1] * d[0]
    d3 = amb[0] * cmd[1] - amb[1] * cmd[0]
    return (d1 * cmd - d2 * amb) / d3


def fit2pointcylinder(p, n, params) -> Optional[Shape]:
    """cylinder.jl:34-125."""
    par = params["common"]["parallelthrdeg"]
    if abs(dot3(n[0], n[1])) > cosd(par):  # Q8: raw normals
        return None
    with np.errstate(invalid="ignore", divide="ignore"):
        an = normalize3(cross3(n[0], n[1]))
        xax = normalize3(_project2plane(an, p[0]))
        yax = normalize3(cross3(an, xax))
        p11 = _projectto2d(xax, yax, an, _project2plane(an, p[0]))
        p12 = _projectto2d(xax, yax, an, _project2plane(an, p[0] + n[0]))
        p21 = _projectto2d(xax, yax, an, _project2plane(an, p[1]))
        p22 = _projectto2d(xax, yax, an, _project2plane(an, p[1] + n[1]))
        ic = _lineintersection(p11, p12, p21, p22)
        c = ic[0] * xax + ic[1] * yax
        nn = [norm3(pt - c - an * dot3(an, pt - c)) for pt in (p[0], p[1])]
        R = (nn[0] + nn[1]) / 2
    return Shape(CYLINDER, an, c, float(R), True)


def fit_cylinder(p, n, params) -> Optional[Shape]:
    """cylinder.jl:135-168."""
    p, n = _v(p), _v(n)
    eps = params["cylinder"]["eps"]
    alpha = params["cylinder"]["alpha"]
    pl = len(p)
    assert pl == len(n) and pl > 2
    fc = fit2pointcylinder(p, n, params)
    if fc is None:
        return None
    thr = math.cos(alpha)
    vert_ok = np.zeros(pl, bool)
    norm_ok = np.zeros(pl, bool)
    inv_ok = np.zeros(pl, bool)
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(pl):
            cn = p[i] - fc.a * dot3(fc.a, p[i] - fc.b) - fc.b
            vert_ok[i] = abs(norm3(cn) - fc.s) < eps
            d = dot3(normalize3(cn), n[i])
            norm_ok[i] = d > thr
            inv_ok[i] = d < -thr
    if not vert_ok.all():
        return None
    if norm_ok.all():
        fc.outwards = True
        return fc
    if inv_ok.all():
        fc.outwards = False
        return fc
    return None


def compatibles_cylinder(sh: Shape, pts, nrm, params) -> np.ndarray:
    """cylinder.jl:194-221."""
    eps = params["cylinder"]["eps"]
    thr = math.cos(params["cylinder"]["alpha"])
    a, c, R = sh.a, sh.b, sh.s
    with np.errstate(invalid="ignore", divide="ignore"):
        h = dot3(a, pts - c)
        cn = pts - a * h[..., None] - c
        ok_r = np.abs(norm3(cn) - R) < eps
        u = normalize3(cn)
        if not sh.outwards:
            u = -u
        return ok_r & (dot3(u, nrm) > thr)


# ---- cone ----------------------------------------------------------------------------


def _rank_julia(m: np.ndarray) -> int:
    """LinearAlgebra.rank: count(s .> min(size)*eps*maximum(s))."""
    if not np.isfinite(m).all():
        return -1
    s = np.linalg.svd(m, compute_uv=False)
    tol = min(m.shape) * np.finfo(F).eps * s.max()
    return int((s > tol).sum())


def fit3pointcone(p, n) -> Optional[Shape]:
    """cone.jl:39-61."""
    r = np.array([[n[i][j] for j in range(3)] for i in range(3)], dtype=F)
    if _rank_julia(r) != 3:
        return None
    ds = np.array([dot3(p[i], n[i]) for i in range(3)])
    rv = np.hstack([r, (-1 * ds)[:, None]])
    if _rank_julia(rv) != 3:
        return None
    ap = np.linalg.solve(r, ds)
    with np.errstate(invalid="ignore", divide="ignore"):
        ax3 = [ap + ((p[i] - ap) / norm3(p[i] - ap)) for i in range(3)]
        ax = normalize3(cross3(ax3[1] - ax3[0], ax3[2] - ax3[0]))
        midp = (ax3[0] + ax3[1] + ax3[2]) / 3
        dirv = normalize3(midp - ap)
        if dot3(ax, dirv) < 0:
            ax = -1 * ax
        angles = [math.acos(_clamp(dot3(normalize3(p[i] - ap), ax))) for i in range(3)]
    op = 2 * (angles[0] + angles[1] + angles[2]) / 3
    return Shape(CONE, ap, ax, float(op), True)


def _clamp(x):
    if x != x:
        return x
    return min(max(x, -1.0), 1.0)


def crossprodtensor(v) -> np.ndarray:
    """utilities.jl:50-54."""
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def pluscrossprod(A: np.ndarray, value, v) -> np.ndarray:
    """pluscrossprod! (utilities.jl:32-43), in place; A is (..., 3, 3), v is (..., 3), value a scalar."""
    A[..., 0, 1] -= value * v[..., 2]
    A[..., 0, 2] += value * v[..., 1]
    A[..., 1, 0] += value * v[..., 2]
    A[..., 1, 2] -= value * v[..., 0]
    A[..., 2, 0] -= value * v[..., 1]
    A[..., 2, 1] += value * v[..., 0]
    return A


def rodrigues(nv: np.ndarray, ct: float, st: float) -> np.ndarray:
    """rodrigues(nv, theta) (utilities.jl:6-11,19-24) for a batch of unit axes nv (n,3), given
    cos(theta), sin(theta): nv nv' + cos (I - nv nv'), then pluscrossprod!(R, sin, nv)."""
    outer = nv[..., :, None] * nv[..., None, :]
    R = outer + ct * (np.eye(3) - outer)
    return pluscrossprod(R, st, nv)


def push2candidatesandlevels(candidates: list, candidate, levels: list, current_level) -> None:
    """utilities.jl:473-481: a fit may return one candidate or an array of candidates."""
    if isinstance(candidate, (list, tuple)):
        candidates.extend(candidate)
        levels.extend([current_level] * len(candidate))
    else:
        candidates.append(candidate)
        levels.append(current_level)


def project2cone(sh: Shape, pts):
    """cone.jl:68-85, vectorised over points; returns (dist, normal)."""
    apex, axis, opang = sh.a, sh.b, sh.s
    pts = np.atleast_2d(pts)
    with np.errstate(invalid="ignore", divide="ignore"):
        to_point = apex - pts
        to_pointn = normalize3(to_point)
        rot_ax = normalize3(cross3(np.broadcast_to(axis, to_pointn.shape), to_pointn))
        comp_n = normalize3(cross3(np.broadcast_to(axis, rot_ax.shape), rot_ax))
        # rodriguesrad(rot_ax, -opang/2): utilities.jl:61-64 -> :19-24 -> :32-43
        nv = normalize3(rot_ax)
        th = -opang / 2
        # (Julia's cos(Inf) throws a DomainError; a non-finite opening angle is treated as NaN = matches nothing)
        ct, st = float(np.cos(th)), float(np.sin(th))
        R = rodrigues(nv, ct, st)
        rv = np.stack(
            [
                R[:, i, 0] * comp_n[:, 0] + R[:, i, 1] * comp_n[:, 1] + R[:, i, 2] * comp_n[:, 2]
                for i in range(3)
            ],
            axis=-1,
        )
        cur_n = normalize3(rv)
        dist = dot3(-cur_n, -to_point)
    return dist, cur_n


def validatecone(sh: Shape, ps, ns, params) -> Optional[Shape]:
    """cone.jl:87-115."""
    eps = params["cone"]["eps"]
    alpha = params["cone"]["alpha"]
    minop = params["cone"]["minconeopang"]
    dist, nr = project2cone(sh, ps)
    for i in range(len(ps)):
        if dist[i] > eps:  # Q6: signed
            return None
    if sh.s < minop:
        return None
    thr = math.cos(alpha)
    with np.errstate(invalid="ignore"):
        d = dot3(nr, ns)
        norm_ok = d > thr
        inv_ok = d < -thr
    if norm_ok.all():
        sh.outwards = True
        return sh
    if inv_ok.all():
        sh.outwards = False
        return sh
    return None


def fit_cone(p, n, params) -> Optional[Shape]:
    """cone.jl:123-128."""
    p, n = _v(p), _v(n)
    fc = fit3pointcone(p, n)
    if fc is None:
        return None
    return validatecone(fc, p, n, params)


def compatibles_cone(sh: Shape, pts, nrm, params) -> np.ndarray:
    """cone.jl:132-153."""
    eps = params["cone"]["eps"]
    thr = math.cos(params["cone"]["alpha"])
    dist, nr = project2cone(sh, pts)
    if not sh.outwards:
        nr = -nr
    with np.errstate(invalid="ignore"):
        return (dot3(nr, nrm) > thr) & (np.abs(dist) < eps)


FIT = {PLANE: fit_plane, SPHERE: fit_sphere, CYLINDER: fit_cylinder, CONE: fit_cone}
COMPAT = {
    PLANE: compatibles_plane,
    SPHERE: compatibles_sphere,
    CYLINDER: compatibles_cylinder,
    CONE: compatibles_cone,
}


def compatibles(sh: Shape, pts, nrm, params, chunk: int = 1 << 18) -> np.ndarray:
    """compatibles* for any shape, chunked so the cone's temporaries stay small."""
    pts = np.asarray(pts, dtype=F)
    nrm = np.asarray(nrm, dtype=F)
    out = np.empty(len(pts), bool)
    f = COMPAT[sh.kind]
    for s in range(0, len(pts), chunk):
        out[s : s + chunk] = f(sh, pts[s : s + chunk], nrm[s : s + chunk], params)
    return out


# --------------------------------------------------------------------------------------
# confidence intervals (confidenceintervals.jl)
# --------------------------------------------------------------------------------------


@dataclass
class ConfidenceInterval:
    min: float
    max: float
    E: float = field(init=False)

    def __post_init__(self):
        if self.min > self.max:
            raise ValueError("out of order")  # confidenceintervals.jl:5
        self.min = float(self.min)
        self.max = float(self.max)
        self.E = (self.min + self.max) / 2


def notsoconfident(x, y) -> ConfidenceInterval:
    # Julia's min/max propagate NaN; Python's do not
    if x != x or y != y:
        ci = ConfidenceInterval.__new__(ConfidenceInterval)
        ci.min = ci.max = ci.E = float("nan")
        return ci
    return ConfidenceInterval(min(x, y), max(x, y))


def isoverlap(i1: ConfidenceInterval, i2: ConfidenceInterval) -> bool:
    if i1.min == i2.min:
        return True
    if i1.min < i2.min:
        return i2.min <= i1.max
    return isoverlap(i2, i1)


def _wrap64(x: int) -> int:
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


def hypergeomdev(N: int, x: int, n: int):
    """confidenceintervals.jl:53-59 with Int64 wrap-around (Q9)."""
    prod = _wrap64(_wrap64(_wrap64(x * n) * (N - x)) * (N - n))
    sq_ = prod / (N - 1)
    sq = 0.0 if sq_ < 0 else math.sqrt(sq_)
    xn = _wrap64(x * n)
    return (xn + sq) / N, (xn - sq) / N


def estimatescore(S1length: int, Plength: int, sigma: int) -> ConfidenceInterval:
    """confidenceintervals.jl:71-74."""
    gmin, gmax = hypergeomdev(-2 - S1length, -2 - Plength, -1 - sigma)
    return notsoconfident(-1 - gmin, -1 - gmax)


def estimatescore_f64(Slength: int, Plength: int, sigma: int) -> ConfidenceInterval:
    """confidenceintervals.jl:53-74 with the products taken in float64, left to right (no Int64 wrap, Q9):
    the interval the formula means.  Used by progressive scoring, whose decisions read
    min/max; the reference's loop only ever reads E, which the wrap does not touch."""
    Np, x, n = float(-2 - Slength), float(-2 - Plength), float(-1 - sigma)
    sq_ = ((x * n) * (Np - x)) * (Np - n) / (Np - 1.0)
    sq = 0.0 if sq_ < 0 else math.sqrt(sq_)
    return notsoconfident(-1 - (x * n + sq) / Np, -1 - (x * n - sq) / Np)


def refine_progressive(pc: "Cloud", params, shapes, scores, evaluated, trace=None) -> None:
    """Progressive subset scoring (iterations.jl:110 "TODO: refine if best.overlap",
    docs/src/ransac.md:137-141, Schnabel et al. 2007 sec. 4.5.1).  `evaluated[i] = [j, sigma, M]`:
    candidate i has been scored on subsets 1..j, with sigma compatible points among their M points.
    While the interval of the best candidate (largest E, first wins) overlaps another one (isoverlap,
    confidenceintervals.jl:29-36), the least-evaluated candidates among the best and its overlappers are
    scored on their next subset and their interval is re-estimated from the union (sigma, M summed).
    Stops when the best stands alone or those candidates have used every subset."""
    r = len(pc.subsets)
    while len(shapes) > 1:
        best, overlap = findhighestscore(scores)
        if not overlap:
            return
        group = [i for i in range(len(scores)) if i == best or isoverlap(scores[i], scores[best])]
        lmin = min(evaluated[i][0] for i in group)
        if lmin >= r:
            return
        for i in group:
            if evaluated[i][0] != lmin:
                continue
            sub = pc.subsets[lmin]  # 0-based id of subset lmin+1
            cp = compatibles(shapes[i], pc.vertices[sub], pc.normals[sub], params)
            if shapes[i].kind != SPHERE:  # same per-type policy as scorecandidate (Q4)
                cp = cp & pc.isenabled[sub]
            evaluated[i][0] += 1
            evaluated[i][1] += int(cp.sum())
            evaluated[i][2] += len(sub)
            scores[i] = estimatescore_f64(evaluated[i][2], pc.size, evaluated[i][1])
            if trace is not None:
                trace.evals += len(sub)
                trace.refined += 1


def prob(n, s, N, k):
    """utilities.jl:262."""
    return 1 - (1 - (n / N) ** k) ** s


# --------------------------------------------------------------------------------------
# cloud, scoring, refit (octree.jl:37-138, shapes/*.jl scorecandidate/refit)
# --------------------------------------------------------------------------------------


@dataclass
class Cloud:
    vertices: np.ndarray  # (N,3) float64
    normals: np.ndarray  # (N,3) float64
    subsets: List[np.ndarray]  # 0-based index arrays (a partition)
    isenabled: np.ndarray = None  # (N,) bool
    size: int = 0

    def __post_init__(self):
        self.vertices = np.asarray(self.vertices, dtype=F)
        self.normals = np.asarray(self.normals, dtype=F)
        self.size = len(self.vertices)
        if self.isenabled is None:
            self.isenabled = np.ones(self.size, bool)


def make_subsets(n: int, numofsubsets: int, perm: np.ndarray) -> List[np.ndarray]:
    """octree.jl:129-135 given the permutation (Julia's randperm stream is not reproducible)."""
    ssl = n // numofsubsets
    subs = [perm[i * ssl : (i + 1) * ssl] for i in range(numofsubsets - 1)]
    subs.append(perm[(numofsubsets - 1) * ssl :])
    return subs


def scorecandidate(pc: Cloud, sh: Shape, subset_id: int, params) -> Tuple[ConfidenceInterval, np.ndarray]:
    sub = pc.subsets[subset_id]
    cp = compatibles(sh, pc.vertices[sub], pc.normals[sub], params)
    if sh.kind != SPHERE:  # Q4: sphere.jl:121,131 never applies `ens`
        cp = cp & pc.isenabled[sub]
    inpoints = sub[cp]
    return estimatescore(len(sub), pc.size, len(inpoints)), inpoints


def refit(sh: Shape, pc: Cloud, params) -> np.ndarray:
    """refit: compatible points among the enabled points of the whole cloud (ascending)."""
    en = np.flatnonzero(pc.isenabled)
    cp = compatibles(sh, pc.vertices[en], pc.normals[en], params)
    return en[cp]


def findhighestscore(scores: Sequence[ConfidenceInterval]) -> Tuple[int, bool]:
    """fitting.jl:140-158 (0-based index, -1 when empty)."""
    if len(scores) == 0:
        return -1, False
    ind = 0
    highest = scores[0].E
    for i, sc in enumerate(scores):
        if sc.E > highest:
            highest = sc.E
            ind = i
    for i, sc in enumerate(scores):
        if i == ind:
            continue
        if isoverlap(sc, scores[ind]):
            return ind, True
    return ind, False


# --------------------------------------------------------------------------------------
# least-squares refit (extension).  The reference's refit keeps the candidate as it
# is ("In our implementation least-square fitting is not used", docs/src/ransac.md:163-169); the
# paper refits it to all compatible points within 3 eps before the extraction.  Definition shared
# with the CUDA library (rsc_lsq.cu): the point set is selected ONCE with the candidate
# (compatibles* with eps scaled by `band`, enabled points only); planes get the total-least-squares
# plane of that set (centroid + smallest-eigenvalue direction of the scatter matrix), the other
# shapes a Levenberg-Marquardt minimisation of the sum of squared point-to-surface distances.
# --------------------------------------------------------------------------------------

LSQ_NPAR = {PLANE: 3, SPHERE: 4, CYLINDER: 7, CONE: 7}
LSQ_MAXIT = 12
LSQ_REL = 1e-12


def lsq_pack(sh: Shape) -> np.ndarray:
    if sh.kind == PLANE:
        return np.array([*sh.a, *sh.b], F)
    if sh.kind == SPHERE:
        return np.array([*sh.a, sh.s], F)
    if sh.kind == CYLINDER:  # axis, center, radius
        return np.array([*sh.a, *sh.b, sh.s], F)
    return np.array([*sh.a, *sh.b, sh.s / 2], F)  # apex, axis, HALF opening angle


def lsq_unpack(kind: int, outwards: bool, x: np.ndarray) -> Shape:
    if kind == PLANE:
        return Shape(PLANE, x[0:3].copy(), x[3:6].copy(), 0.0, True)
    if kind == SPHERE:
        return Shape(SPHERE, x[0:3].copy(), np.zeros(3), float(x[3]), outwards)
    if kind == CYLINDER:
        return Shape(CYLINDER, x[0:3].copy(), x[3:6].copy(), float(x[6]), outwards)
    return Shape(CONE, x[0:3].copy(), x[3:6].copy(), float(2 * x[6]), outwards)


def lsq_normalise(kind: int, x: np.ndarray) -> np.ndarray:
    """bring a stepped parameter vector back to the reference's conventions: unit axis; cylinder
    centre on the plane through the origin perpendicular to the axis (cylinder.jl:114)"""
    x = x.copy()
    if kind == CYLINDER:
        a = x[0:3] / math.sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2])
        c = x[3:6]
        x[0:3] = a
        x[3:6] = c - a * (a[0] * c[0] + a[1] * c[1] + a[2] * c[2])
    elif kind == CONE:
        x[3:6] = x[3:6] / math.sqrt(x[3] * x[3] + x[4] * x[4] + x[5] * x[5])
    return x


def lsq_residual_jacobian(kind: int, x: np.ndarray, P: np.ndarray):
    """residual r (n,) and Jacobian J (n, npar) of the signed point-to-surface distance"""
    if kind == PLANE:  # moments about the old point: "J" = v, "r" = 1 (A = sum v v^T, g = sum v, cost = n)
        v = P - x[0:3]
        return np.ones(len(P)), v
    if kind == SPHERE:
        v = P - x[0:3]
        rho = np.sqrt(v[:, 0] * v[:, 0] + v[:, 1] * v[:, 1] + v[:, 2] * v[:, 2])
        u = v / rho[:, None]
        return rho - x[3], np.c_[-u, -np.ones(len(P))]
    if kind == CYLINDER:
        a, c, R = x[0:3], x[3:6], x[6]
        v = P - c
        h = v[:, 0] * a[0] + v[:, 1] * a[1] + v[:, 2] * a[2]
        w = v - h[:, None] * a
        rho = np.sqrt(w[:, 0] * w[:, 0] + w[:, 1] * w[:, 1] + w[:, 2] * w[:, 2])
        wh = w / rho[:, None]
        return rho - R, np.c_[-h[:, None] * wh, -wh, -np.ones(len(P))]
    ap, a, th = x[0:3], x[3:6], x[6]
    sn, cs = math.sin(th), math.cos(th)
    v = P - ap
    h = v[:, 0] * a[0] + v[:, 1] * a[1] + v[:, 2] * a[2]
    w = v - h[:, None] * a
    rho = np.sqrt(w[:, 0] * w[:, 0] + w[:, 1] * w[:, 1] + w[:, 2] * w[:, 2])
    wh = w / rho[:, None]
    dv = sn * a[None, :] - cs * wh  # d r / d v
    return h * sn - rho * cs, np.c_[-dv, sn * v + (cs * h)[:, None] * wh, h * cs + rho * sn]


def lsq_accumulate(kind: int, x: np.ndarray, P: np.ndarray):
    """normal equations of one pass: A = J^T J, g = J^T r, cost = r^T r (points with a non-finite
    residual or Jacobian -- on the axis, at the centre -- are left out)"""
    with np.errstate(invalid="ignore", divide="ignore"):
        r, J = lsq_residual_jacobian(kind, x, P)
    ok = np.isfinite(r) & np.isfinite(J).all(axis=1)
    r, J = r[ok], J[ok]
    return J.T @ J, J.T @ r, float(r @ r)


def lsq_cholesky_solve(A: np.ndarray, b: np.ndarray) -> Optional[np.ndarray]:
    """solve A x = b for a symmetric positive definite A (None if a pivot is not positive)"""
    n = len(b)
    L = np.zeros((n, n))
    for j in range(n):
        d = A[j, j] - sum(L[j, k] * L[j, k] for k in range(j))
        if not (d > 0) or not math.isfinite(d):
            return None
        L[j, j] = math.sqrt(d)
        for i in range(j + 1, n):
            L[i, j] = (A[i, j] - sum(L[i, k] * L[j, k] for k in range(j))) / L[j, j]
    y = np.zeros(n)
    for i in range(n):
        y[i] = (b[i] - sum(L[i, k] * y[k] for k in range(i))) / L[i, i]
    xs = np.zeros(n)
    for i in reversed(range(n)):
        xs[i] = (y[i] - sum(L[k, i] * xs[k] for k in range(i + 1, n))) / L[i, i]
    return xs


def lsq_smallest_eigvec3(M: np.ndarray) -> np.ndarray:
    """eigenvector of the smallest eigenvalue of a symmetric 3x3 matrix: cyclic Jacobi, fixed sweep
    order (0,1), (0,2), (1,2), 30 sweeps at most"""
    A = M.astype(F).copy()
    V = np.eye(3)
    for _ in range(30):
        off = abs(A[0, 1]) + abs(A[0, 2]) + abs(A[1, 2])
        if off <= 1e-300 or off <= 1e-17 * (abs(A[0, 0]) + abs(A[1, 1]) + abs(A[2, 2])):
            break
        for p_, q_ in ((0, 1), (0, 2), (1, 2)):
            if A[p_, q_] == 0.0:
                continue
            tau = (A[q_, q_] - A[p_, p_]) / (2 * A[p_, q_])
            t = (1.0 if tau >= 0 else -1.0) / (abs(tau) + math.sqrt(1 + tau * tau))
            c = 1 / math.sqrt(1 + t * t)
            s_ = t * c
            Jm = np.eye(3)
            Jm[p_, p_] = Jm[q_, q_] = c
            Jm[p_, q_], Jm[q_, p_] = s_, -s_
            A = Jm.T @ A @ Jm
            V = V @ Jm
    k = int(np.argmin([A[0, 0], A[1, 1], A[2, 2]]))
    return V[:, k].copy()


def lsq_select(sh: Shape, pc: "Cloud", params, band: float = 3.0) -> np.ndarray:
    """indices of the enabled points compatible with `sh` inside band * eps (and alpha)"""
    name = SHAPE_NAMES[sh.kind]
    p3 = {k: dict(v) for k, v in params.items()}
    p3[name]["eps"] = params[name]["eps"] * band
    en = np.flatnonzero(pc.isenabled)
    return en[compatibles(sh, pc.vertices[en], pc.normals[en], p3)]


def lsq_refine(sh: Shape, pc: "Cloud", params, band: float = 3.0):
    """-> (refined shape, number of points used, rms distance).  The candidate is returned unchanged
    if fewer than 2 * npar points are selected or the minimisation leaves the finite numbers."""
    sel = lsq_select(sh, pc, params, band)
    npar = LSQ_NPAR[sh.kind]
    n = len(sel)
    if n < 2 * npar:
        return sh, n, float("nan")
    P = pc.vertices[sel]
    x = lsq_pack(sh)
    if sh.kind == PLANE:
        A, g, cnt = lsq_accumulate(PLANE, x, P)
        mean = g / cnt
        M = A - cnt * np.outer(mean, mean)
        nv = lsq_smallest_eigvec3(M)
        nv = nv / math.sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2])
        if nv[0] * x[3] + nv[1] * x[4] + nv[2] * x[5] < 0:
            nv = -nv
        out = np.array([*(x[0:3] + mean), *nv])
        if not np.isfinite(out).all():
            return sh, n, float("nan")
        rms = math.sqrt(max(0.0, float(nv @ M @ nv)) / cnt)
        return lsq_unpack(PLANE, True, out), n, rms
    x = lsq_normalise(sh.kind, x)
    A, g, cost = lsq_accumulate(sh.kind, x, P)
    lam = 1e-3
    for _ in range(LSQ_MAXIT):
        D = A + lam * np.diag(np.diag(A)) + 1e-12 * np.trace(A) / npar * np.eye(npar)
        d = lsq_cholesky_solve(D, -g)
        if d is None or not np.isfinite(d).all():
            lam *= 10
            continue
        x1 = lsq_normalise(sh.kind, x + d)
        A1, g1, cost1 = lsq_accumulate(sh.kind, x1, P)
        if math.isfinite(cost1) and cost1 <= cost:
            rel = (cost - cost1) / max(cost, 1e-300)
            x, A, g, cost = x1, A1, g1, cost1
            lam = max(lam / 10, 1e-9)
            if rel < LSQ_REL:
                break
        else:
            lam *= 10
    if not np.isfinite(x).all():
        return sh, n, float("nan")
    return lsq_unpack(sh.kind, sh.outwards, x), n, math.sqrt(cost / n)


# --------------------------------------------------------------------------------------
# Philox4x32-10 (Salmon et al., SC'11) -- the stream the CUDA sampler uses, so that the
# oracle and the device draw identical minimal sets (the reference's own RNG stream is
# Julia-version dependent, Q19, and therefore not part of the parity contract).
# --------------------------------------------------------------------------------------

_PH_M0, _PH_M1 = 0xD2511F53, 0xCD9E8D57
_PH_W0, _PH_W1 = 0x9E3779B9, 0xBB67AE85
_M32 = 0xFFFFFFFF


def philox4x32(counter: Tuple[int, int, int, int], key: Tuple[int, int]) -> Tuple[int, int, int, int]:
    c0, c1, c2, c3 = counter
    k0, k1 = key
    for _ in range(10):
        p0 = _PH_M0 * c0
        p1 = _PH_M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _M32, p1 & _M32, ((p0 >> 32) ^ c3 ^ k1) & _M32, p0 & _M32
        k0 = (k0 + _PH_W0) & _M32
        k1 = (k1 + _PH_W1) & _M32
    return c0, c1, c2, c3


class SetStream:
    """Random draws of ONE minimal set: Philox key = seed, counter = (draw/2, set_lo, set_hi, 0).

    Each Philox block yields two 64-bit words; rand_below(n) = mulhi64(word, n).
    """

    def __init__(self, seed: int, set_id: int):
        self.key = (seed & _M32, (seed >> 32) & _M32)
        self.set_id = set_id
        self.ndraw = 0
        self._blk = None

    def next_u64(self) -> int:
        i = self.ndraw
        if i % 2 == 0:
            self._blk = philox4x32((i // 2, self.set_id & _M32, (self.set_id >> 32) & _M32, 0), self.key)
        b = self._blk
        self.ndraw += 1
        return (b[0] | (b[1] << 32)) if i % 2 == 0 else (b[2] | (b[3] << 32))

    def rand_below(self, n: int) -> int:
        return (self.next_u64() * n) >> 64


def sample_minimal_set(pc: Cloud, drawN: int, stream: SetStream, enabled_idx: Optional[np.ndarray] = None):
    """samplepointcloud4! (fitting.jl:383-430) restricted to the root cell (Q1).

    Returns (ok, code, idx[drawN]) -- code 0 = too few enabled points, 1 = duplicate index.
    """
    N = pc.size
    r1 = stream.rand_below(N)
    while not pc.isenabled[r1]:
        r1 = stream.rand_below(N)
    if enabled_idx is None:
        enabled_idx = np.flatnonzero(pc.isenabled)
    ne = len(enabled_idx)
    idx = np.zeros(drawN, np.int64)
    if ne < drawN:
        return False, 0, idx
    idx[0] = r1
    for k in range(1, drawN):
        nexti = stream.rand_below(ne)
        if idx[0] == enabled_idx[nexti]:
            nexti = stream.rand_below(ne)
        idx[k] = enabled_idx[nexti]
    for i in range(1, drawN):
        for j in range(i):
            if idx[i] == idx[j]:
                return False, 1, idx
    return True, 1, idx


# --------------------------------------------------------------------------------------
# Flattened octree + level-weighted cell sampler.  NOT the shipped reference's
# behaviour: there `levelweight`/`levelscore` are swapped at construction (Q1, octree.jl:82-84), every
# minimal set comes from the root cell and the octree never matters.  This restates what the code is
# written to do (octree.jl:158-244, fitting.jl:383-430, octree.jl:198-205) on a Morton-ordered grid
# hierarchy, and is the definition the device sampler is tested against:
#   * root cell = bounding box of the cloud, every level halves each axis (like child_boundary,
#     octree.jl:169); level 1 = root, level l uses the top l-1 bits per axis of the quantised
#     coordinates q = floor((x - lo) / (hi - lo) * 2^D), D = nlevels - 1 (clamped to 2^D - 1);
#   * a cell is split while it holds more than 8 points (OctreeRefinery(8), octree.jl:163-165,241):
#     leafdepth(p) = first level at which p's cell holds <= 8 points, capped at nlevels;
#   * enabled points of a cell are enumerated in Morton order (ties: original index).
# --------------------------------------------------------------------------------------


def smallestdistance(points):
    """utilities.jl:187-199: smallest pairwise distance of a point list (O(n^2); not used by the loop -- its only
    call, iterations.jl:55, is commented out -- restated so that EVERY vector of test/utilitytests.jl pins something)."""
    P = np.asarray(points, dtype=F)
    assert len(P) > 1, "At least two point is needed for that."
    ld = float(np.sqrt(((P[1] - P[0]) ** 2).sum()))
    for i in range(len(P)):
        for j in range(len(P)):
            if i != j:
                d = float(np.sqrt(((P[i] - P[j]) ** 2).sum()))
                ld = d if d < ld else ld
    return ld


def findAABB(points):
    """utilities.jl:125-136: axis-aligned bounding box (min corner, max corner) of a point list; the
    octree's root box (octree.jl:239).  NaN coordinates never replace a bound, like the reference's
    `>` / `<` comparisons."""
    P = np.asarray(points, dtype=F)
    mn, mx = P[0].copy(), P[0].copy()
    for j in range(P.shape[1]):
        col = P[:, j]
        col = col[col == col]
        if len(col):
            mn[j] = min(mn[j], col.min()) if mn[j] == mn[j] else mn[j]
            mx[j] = max(mx[j], col.max()) if mx[j] == mx[j] else mx[j]
    return mn, mx


def iswithinrectangle(vmin, vmax, p) -> bool:
    """octree.jl:187-196: membership test of the reference's octree refinement -- "bottom/left" (the low
    faces) is outside, "top/right" (the high faces) is inside.  (The flattened Morton octree below
    quantises with floor(), i.e. low faces inside: points exactly on a division plane may land in the
    neighbouring cell -- one of the reasons its sampling is an extension, not a parity claim.)"""
    for i in range(3):
        if not (vmin[i] < p[i]):
            return False
        if not (vmax[i] >= p[i]):
            return False
    return True


# --------------------------------------------------------------------------------------
# Parameter-space bitmap + largest connected component (second half).
# The reference carries this as DEAD code (src/parameterspacebitmap.jl is not included by RANSAC.jl,
# its test file test/parameterspacebitmap.jl is commented out in runtests.jl, and `label_components`
# comes from ImageMorphology, which is not a dependency): docs/src/ransac.md:106-112 lists "part of the
# largest connected component in the parameter space bitmap" as the third compatibility criterion and
# says it is not considered.  Restated literally below (bitmapparameters, largestconncomp) and pinned to
# the two known-answer test sets of test/parameterspacebitmap.jl; the extension the CUDA library
# implements (shape_parameters2d + bitmap_filter, rsc_bitmap.cu) follows.
# --------------------------------------------------------------------------------------


def arbitrary_orthogonal(vec) -> np.ndarray:
    """utilities.jl:84-92: a vector orthogonal to `vec` (cross product with the unit vector of its smallest component)."""
    v = normalize3(_v(vec))
    b0 = (v[0] < v[1]) and (v[0] < v[2])
    b1 = (v[1] <= v[0]) and (v[1] < v[2])
    b2 = (v[2] <= v[0]) and (v[2] <= v[1])
    rv = np.array([float(b0), float(b1), float(b2)])
    rv = rv / math.sqrt(float(rv @ rv))
    return cross3(v, rv)


def project2plane(sh: Shape, pts) -> np.ndarray:
    """plane.jl:82-103: coordinates of the points in a frame of the plane (x, y in the plane, z along the normal)."""
    p7 = sh.params7()
    o_z = normalize3(p7[3:6])
    o_x = normalize3(arbitrary_orthogonal(o_z))
    o_y = normalize3(cross3(o_z, o_x))
    V = np.asarray(pts, dtype=F) - p7[0:3]
    return np.stack([V @ o_x, V @ o_y, V @ o_z], axis=1)


def bitmapparameters(parameters, compatibility, beta: float, idsource=None):
    """parameterspacebitmap.jl:12-46, literally (1-based cells kept as 0-based arrays [x-1, y-1]): the box is
    widened by 0.1, a cell keeps the FIRST id projected into it, and places on the border (0, xs, ys) are
    dropped ("boundserror") -- quirks included.  Returns (bitmap (xs, ys) bool, idxmap (xs, ys) int (0 = empty;
    ids as given), (bx, by))."""
    P = np.asarray(parameters, dtype=F)
    comp = np.asarray(compatibility, dtype=bool)
    ids = np.arange(1, len(P) + 1) if idsource is None else np.asarray(idsource)
    assert len(P) == len(comp) == len(ids), "Everything must have the same length."
    miv, mav = findAABB(P[:, :2])
    minv, maxv = miv - 0.1, mav + 0.1
    xs = int(np.round((maxv[0] - minv[0]) / beta))  # Julia's round: half to even, like numpy's
    ys = int(np.round((maxv[1] - minv[1]) / beta))
    assert xs > 0 and ys > 0, "max-min should be positive."
    bx, by = (maxv[0] - minv[0]) / xs, (maxv[1] - minv[1]) / ys
    bitmap = np.zeros((xs, ys), bool)
    idxmap = np.zeros((xs, ys), np.int64)
    for i in range(len(P)):
        if not comp[i]:
            continue
        xp = int(math.ceil((P[i, 0] - minv[0]) / bx))
        yp = int(math.ceil((P[i, 1] - minv[1]) / by))
        if xp != 0 and yp != 0 and xp != xs and yp != ys:
            if not bitmap[xp - 1, yp - 1]:
                bitmap[xp - 1, yp - 1] = True
                idxmap[xp - 1, yp - 1] = ids[i]
    return bitmap, idxmap, (bx, by)


def label_components(bimage: np.ndarray, eight: bool = False) -> np.ndarray:
    """ImageMorphology.label_components as the reference uses it (parameterspacebitmap.jl:68): 0 = background,
    components numbered in the order their first cell is met in COLUMN-MAJOR order (first index fastest),
    4-connectivity (`1:ndims`) or 8-connectivity (`trues(3,3)`).  Plain flood fill."""
    B = np.asarray(bimage, bool)
    xs, ys = B.shape
    lab = np.zeros((xs, ys), np.int64)
    nxt = 0
    nb = [(1, 0), (-1, 0), (0, 1), (0, -1)] + ([(1, 1), (1, -1), (-1, 1), (-1, -1)] if eight else [])
    for y in range(ys):
        for x in range(xs):
            if not B[x, y] or lab[x, y]:
                continue
            nxt += 1
            lab[x, y] = nxt
            stack = [(x, y)]
            while stack:
                cx, cy = stack.pop()
                for dx, dy in nb:
                    ux, uy = cx + dx, cy + dy
                    if 0 <= ux < xs and 0 <= uy < ys and B[ux, uy] and not lab[ux, uy]:
                        lab[ux, uy] = nxt
                        stack.append((ux, uy))
    return lab


def largestconncomp(bimage, indmap, connectivity="default") -> list:
    """parameterspacebitmap.jl:60-109: the ids (`indmap[x][y]` = list of ids of the cell) of the largest
    connected component, cells in column-major order.  connectivity: "default" / range(1, 3) = 4-connected,
    "eight" / a 3x3 all-true array = 8-connected.  The first of equally large components wins (argmax)."""
    if isinstance(connectivity, str):
        if connectivity not in ("default", "eight"):
            raise ValueError(f"No such key implemented: {connectivity}.")
        eight = connectivity == "eight"
    else:
        eight = np.asarray(connectivity).ndim == 2
    B = np.asarray(bimage, bool)
    lab = label_components(B, eight)
    nlab = int(lab.max())
    if nlab == 0:
        raise ValueError("collection must be non-empty")  # the reference's argmax on an empty list (its own TODO)
    sizes = np.bincount(lab.ravel(), minlength=nlab + 1)
    best = int(np.argmax(sizes[1:])) + 1
    inds = []
    xs, ys = B.shape
    for y in range(ys):
        for x in range(xs):
            if lab[x, y] == best:
                cell = indmap[x][y]
                inds.extend(cell if isinstance(cell, (list, tuple, np.ndarray)) else [cell])
    return inds


# ---- the extension: the third compatibility criterion as a filter on an inlier list -----------------
# (definition shared with rsc_bitmap.cu).  2-D parameters of a point on the shape:
#   plane     the first two coordinates of project2plane's frame (plane.jl:82-103) built with stable_orthogonal
#   sphere    (R phi, R sin(lat)) about the z axis -- Lambert's equal-area cylinder; phi wraps
#   cylinder  (R phi, h) in the frame (x, y, axis), x = stable_orthogonal(axis); phi wraps
#   cone      (r_ref phi, s): azimuth about the axis and slant distance from the apex, r_ref = sin(opang/2) times
#             the mid slant distance of the given points (so that cells are ~square there); phi wraps
# Cells: ix = floor((u - umin) / bu), iy = floor((v - vmin) / bv) over the bounding box of the parameters, with
# nu = max(1, round(range / beta)) cells (wrapping axis: nu = max(3, round(2 pi R / beta)) cells over the full
# turn); a component is 4- or 8-connected (x wraps where phi does); the largest one by number of CELLS wins,
# ties go to the component holding the smallest column-major cell index; its points are kept, in input order.


def stable_orthogonal(vec) -> np.ndarray:
    """arbitrary_orthogonal with the smallest component taken by MAGNITUDE: the reference's version (smallest by
    value) returns the zero vector whenever the smallest component is the only non-zero one, e.g. for (0, 0, -1),
    and its project2plane then yields NaNs.  Used by the extension's frames."""
    v = normalize3(_v(vec))
    a = np.abs(v)
    b0 = (a[0] < a[1]) and (a[0] < a[2])
    b1 = (a[1] <= a[0]) and (a[1] < a[2])
    b2 = (a[2] <= a[0]) and (a[2] <= a[1])
    rv = np.array([float(b0), float(b1), float(b2)])
    return cross3(v, rv)


def _dot_rows(V, o):
    """row-wise dot product, summed left to right without contraction (what rsc_bitmap.cu computes)"""
    return (V[:, 0] * o[0] + V[:, 1] * o[1]) + V[:, 2] * o[2]


def shape_parameters2d(sh: Shape, pts):
    """(raw parameters (n, 2), scale of the first one or None, wraps?) -- u = scale * raw[:, 0] where a scale is
    given (cone: it depends on the points), else raw[:, 0]"""
    P = np.asarray(pts, dtype=F)
    p7 = sh.params7()
    if sh.kind == PLANE:
        o_z = normalize3(p7[3:6])
        o_x = normalize3(stable_orthogonal(o_z))
        o_y = normalize3(cross3(o_z, o_x))
        V = P - p7[0:3]
        return np.stack([_dot_rows(V, o_x), _dot_rows(V, o_y)], axis=1), None, False
    if sh.kind == SPHERE:
        R = float(p7[3])
        V = P - p7[0:3]
        r = np.sqrt(_dot_rows(V, V.T))
        phi = np.arctan2(V[:, 1], V[:, 0])
        return np.stack([R * phi, R * (V[:, 2] / r)], axis=1), None, True
    if sh.kind == CYLINDER:
        a = normalize3(p7[0:3])
        ox = normalize3(stable_orthogonal(a))
        oy = normalize3(cross3(a, ox))
        V = P - p7[3:6]
        R = float(p7[6])
        phi = np.arctan2(_dot_rows(V, oy), _dot_rows(V, ox))
        return np.stack([R * phi, _dot_rows(V, a)], axis=1), None, True
    a = normalize3(p7[3:6])
    ox = normalize3(stable_orthogonal(a))
    oy = normalize3(cross3(a, ox))
    V = P - p7[0:3]
    s = np.sqrt(_dot_rows(V, V.T))
    phi = np.arctan2(_dot_rows(V, oy), _dot_rows(V, ox))
    r_ref = abs(math.sin(float(p7[6]) / 2)) * (0.5 * (float(s.min()) + float(s.max()))) if len(s) else 1.0
    r_ref = max(r_ref, 1e-300)
    return np.stack([phi, s], axis=1), r_ref, True


def bitmap_period(sh: Shape, scale) -> float:
    """length of the wrapping axis: the full turn"""
    p7 = sh.params7()
    if sh.kind == SPHERE:
        return 2 * math.pi * float(p7[3])
    if sh.kind == CYLINDER:
        return 2 * math.pi * float(p7[6])
    return 2 * math.pi * scale


def bitmap_filter(sh: Shape, pts, beta: float, eight: bool = False):
    """indices (into pts, ascending) of the points whose parameter-space cell belongs to the largest connected
    component; also returns (nu, nv, number of components, cells of the largest)."""
    P = np.asarray(pts, dtype=F)
    if len(P) == 0:
        return np.zeros(0, np.int64), (0, 0, 0, 0)
    uv, scale, wraps = shape_parameters2d(sh, P)
    if scale is not None:
        uv = np.stack([scale * uv[:, 0], uv[:, 1]], axis=1)
    period = bitmap_period(sh, scale) if wraps else None
    vmin, vmax = float(uv[:, 1].min()), float(uv[:, 1].max())
    nv = max(1, int(np.round((vmax - vmin) / beta)))
    bv = (vmax - vmin) / nv if vmax > vmin else 1.0
    if period is None:
        umin, umax = float(uv[:, 0].min()), float(uv[:, 0].max())
        nu = max(1, int(np.round((umax - umin) / beta)))
        bu = (umax - umin) / nu if umax > umin else 1.0
    else:
        umin = -period / 2
        nu = max(3, int(np.round(period / beta)))
        bu = period / nu
    ix = np.minimum(np.floor((uv[:, 0] - umin) / bu).astype(np.int64), nu - 1)
    iy = np.minimum(np.floor((uv[:, 1] - vmin) / bv).astype(np.int64), nv - 1)
    ix = np.maximum(ix, 0)
    B = np.zeros((nu, nv), bool)
    B[ix, iy] = True
    # union-find free labelling with optional wrap in x: flood fill
    lab = np.full((nu, nv), -1, np.int64)
    nb = [(1, 0), (-1, 0), (0, 1), (0, -1)] + ([(1, 1), (1, -1), (-1, 1), (-1, -1)] if eight else [])
    sizes, roots = [], []
    for y in range(nv):
        for x in range(nu):
            if not B[x, y] or lab[x, y] >= 0:
                continue
            k = len(sizes)
            lab[x, y] = k
            cnt, stack = 0, [(x, y)]
            while stack:
                cx, cy = stack.pop()
                cnt += 1
                for dx, dy in nb:
                    ux, uy = cx + dx, cy + dy
                    if period is not None:
                        ux %= nu
                    if 0 <= ux < nu and 0 <= uy < nv and B[ux, uy] and lab[ux, uy] < 0:
                        lab[ux, uy] = k
                        stack.append((ux, uy))
            sizes.append(cnt)
            roots.append(x + nu * y)
    best = max(range(len(sizes)), key=lambda k: (sizes[k], -roots[k]))
    keep = np.flatnonzero(lab[ix, iy] == best)
    return keep, (nu, nv, len(sizes), sizes[best])


def morton3(q: np.ndarray, D: int) -> np.ndarray:
    """interleave the D low bits of q[:,0], q[:,1], q[:,2] (x is the most significant of each triple)"""
    code = np.zeros(len(q), np.int64)
    for b in range(D - 1, -1, -1):
        code = (code << 3) | (((q[:, 0] >> b) & 1) << 2) | (((q[:, 1] >> b) & 1) << 1) | ((q[:, 2] >> b) & 1)
    return code


class MortonOctree:
    def __init__(self, vertices, nlevels: int):
        assert 1 <= nlevels <= 11
        V = np.asarray(vertices, dtype=F)
        self.nlevels = nlevels
        D = self.D = nlevels - 1
        lo, hi = findAABB(V)  # octree.jl:239
        w = hi - lo
        w[w == 0] = 1.0
        q = np.floor((V - lo) / w * float(1 << D)).astype(np.int64)
        q = np.clip(q, 0, (1 << D) - 1)
        code = morton3(q, D)
        self.perm = np.argsort(code, kind="stable").astype(np.int64)  # sorted position -> point index
        self.codes = code[self.perm]
        self.inv = np.empty_like(self.perm)
        self.inv[self.perm] = np.arange(len(V))
        self.leafdepth = np.full(len(V), nlevels, np.int32)  # per point (original index)
        undecided = np.ones(len(V), bool)
        for l in range(1, nlevels + 1):
            a, b = self._ranges(code, l)
            small = undecided & ((b - a) <= 8)
            self.leafdepth[small] = l
            undecided &= ~small

    def _ranges(self, code, level):
        shift = 3 * (self.D - (level - 1))
        pre = code >> shift
        a = np.searchsorted(self.codes >> shift, pre, side="left")
        b = np.searchsorted(self.codes >> shift, pre, side="right")
        return a, b

    def cell_range(self, point: int, level: int):
        """[a, b) in Morton order of the level-`level` cell containing `point`"""
        shift = 3 * (self.D - (level - 1))
        pre = self.codes[self.inv[point]] >> shift
        sc = self.codes >> shift
        return int(np.searchsorted(sc, pre, side="left")), int(np.searchsorted(sc, pre, side="right"))


def level_cumsum(levelweight) -> np.ndarray:
    """cum[j-1] = levelweight[1] + ... + levelweight[j], summed left to right in float64"""
    out, c = [], 0.0
    for x in levelweight:
        c += float(x)
        out.append(c)
    return np.array(out)


def draw_level(u64: int, cum: np.ndarray, leafdepth: int) -> int:
    """Level of the cell the other k-1 points come from: drawn from the level distribution restricted
    to 1..leafdepth (docs/src/ransac.md:80-82 -- the shipped code takes argmax(levelweight[1:leafdepth])
    instead, fitting.jl:401, which never leaves level 1).  1-based."""
    target = (float(u64) * 2.0 ** -64) * float(cum[leafdepth - 1])
    for l in range(1, leafdepth + 1):
        if cum[l - 1] > target:
            return l
    return leafdepth


def sample_minimal_set_octree(pc: "Cloud", oct: MortonOctree, drawN: int, stream: SetStream, cum, en_sorted: np.ndarray):
    """samplepointcloud4! (fitting.jl:383-430) on the flattened octree.  `en_sorted` = isenabled in
    Morton order, `cum` = level_cumsum(levelweight).  Returns (ok, level, idx)."""
    N = pc.size
    r1 = stream.rand_below(N)
    while not pc.isenabled[r1]:
        r1 = stream.rand_below(N)
    level = draw_level(stream.next_u64(), cum, int(oct.leafdepth[r1]))
    a, b = oct.cell_range(r1, level)
    pos = np.flatnonzero(en_sorted[a:b]) + a
    ne = len(pos)
    idx = np.zeros(drawN, np.int64)
    if ne < drawN:
        return False, level, idx
    idx[0] = r1
    for k in range(1, drawN):
        nexti = stream.rand_below(ne)
        if idx[0] == oct.perm[pos[nexti]]:
            nexti = stream.rand_below(ne)
        idx[k] = oct.perm[pos[nexti]]
    for i in range(1, drawN):
        for j in range(i):
            if idx[i] == idx[j]:
                return False, level, idx
    return True, level, idx


def updatelevelweight(levelweight: np.ndarray, levelscore: np.ndarray) -> np.ndarray:
    """octree.jl:198-205 with its default x = 9//10; unchanged while no level has a score yet (w == 0 would give NaN).
    x is a Rational in the reference: `x*σ[i]` promotes it to Float64(9//10) = 0.9, but `(1-x)/length(P)` is the exact
    rational 1//(10 n), converted only when it is added to the float term -- the correctly rounded 1/(10 n), whereas
    the float expression (1 - 0.9)/n would be one ulp off."""
    P, sg = levelweight, levelscore
    w = 0.0
    for i in range(len(P)):
        w += sg[i] / P[i]
    if not w > 0.0:
        return P.copy()
    floor_w = 1.0 / (10.0 * len(P))
    return np.array([0.9 * sg[i] / (w * P[i]) + floor_w for i in range(len(P))])


# --------------------------------------------------------------------------------------
# the loop (iterations.jl:35-162)
# --------------------------------------------------------------------------------------


@dataclass
class Extracted:
    shape: Shape
    inpoints: np.ndarray


@dataclass
class RansacTrace:
    iterations: int = 0
    sets_drawn: int = 0
    candidates_scored: int = 0
    evals: int = 0
    extracted_at: List[int] = field(default_factory=list)
    levelweight: Optional[np.ndarray] = None
    refined: int = 0  # progressive scoring: (candidate, subset) evaluations beyond subset 1


def forcefit(p, n, params) -> List[Shape]:
    """forcefitshapes! (fitting.jl:165-173): candidates in shape_types order."""
    out = []
    for s in params["iteration"]["shape_types"]:
        sh = FIT[s](p, n, params)
        if sh is not None:
            out.append(sh)
    return out


def ransac(
    pc: Cloud,
    params: dict,
    setenabled: bool = True,
    seed: int = 1234,
    minimal_sets: Optional[Callable[[int, int], Optional[np.ndarray]]] = None,
    trace: Optional[RansacTrace] = None,
    octree: Optional[MortonOctree] = None,
    progressive: bool = False,
    lsq: bool = False,
    bitmap: Optional[Tuple[float, bool]] = None,
    lw_period: int = 1,
) -> List[Extracted]:
    """iterations.jl:14-21 + :35-162.

    `progressive=True` (extension) refines overlapping scores on further subsets before
    the extraction test (refine_progressive); intervals are then the float64 ones (estimatescore_f64).
    `lsq=True` (extension) refits the best candidate by least squares before extracting it.
    `bitmap=(beta, eight)` (extension) extracts only the compatible points in the largest connected
    component of the shape's parameter-space bitmap (bitmap_filter; docs/src/ransac.md:106-112).

    `minimal_sets(k, i)` may supply the index triple of minimal set i of iteration k (or None
    for a failed sample); by default the Philox sampler above is used with set_id =
    (k-1)*minsubsetN + i.
    """
    it = params["iteration"]
    drawN, minsubsetN, prob_det, tau = it["drawN"], it["minsubsetN"], it["prob_det"], it["tau"]
    itermax = it["itermax"]
    sidx = {"lengthC": 0, "allcand": 1, "nofminset": 2}
    if setenabled:
        pc.isenabled[:] = True
    shapes: List[Shape] = []
    scores: List[ConfidenceInterval] = []
    inpts: List[np.ndarray] = []
    evaluated: List[list] = []  # progressive scoring state per stored candidate
    extracted: List[Extracted] = []
    cc = [0, 0, 0]
    tr = trace if trace is not None else RansacTrace()
    if octree is not None:  # level-weighted cell sampler (8(f)-1); un-swapped initial values (octree.jl:82-83)
        levelweight = np.full(octree.nlevels, 1.0 / octree.nlevels)
        levelscore = np.zeros(octree.nlevels)
    for k in range(1, itermax + 1):
        if int(pc.isenabled.sum()) < tau:
            break
        tr.iterations = k
        cands: List[Shape] = []
        levels: List[int] = []
        en_idx = np.flatnonzero(pc.isenabled)
        if octree is not None:
            cum = level_cumsum(levelweight)
            en_sorted = pc.isenabled[octree.perm]
        for i in range(minsubsetN):
            if minimal_sets is not None:
                sd = minimal_sets(k, i)
                if sd is None:
                    continue
            elif octree is not None:
                ok, lvl, sd = sample_minimal_set_octree(pc, octree, drawN, SetStream(seed, (k - 1) * minsubsetN + i), cum, en_sorted)
                if not ok:
                    continue
                for fitted in forcefit(pc.vertices[sd], pc.normals[sd], params):  # fitting.jl:165-173
                    push2candidatesandlevels(cands, fitted, levels, lvl)
                continue
            else:
                ok, _, sd = sample_minimal_set(pc, drawN, SetStream(seed, (k - 1) * minsubsetN + i), en_idx)
                if not ok:
                    continue
            cands.extend(forcefit(pc.vertices[sd], pc.normals[sd], params))
        cc[1] += len(cands)
        lv_n = np.zeros(octree.nlevels if octree is not None else 0, np.int64)
        lv_s = np.zeros_like(lv_n)
        for ci, c in enumerate(cands):  # scorecandidates! (fitting.jl:181-190), subset 1 only (Q10)
            sc, ip = scorecandidate(pc, c, 0, params)
            if progressive:
                sc = estimatescore_f64(len(pc.subsets[0]), pc.size, len(ip))
            shapes.append(c)
            scores.append(sc)
            inpts.append(ip)
            evaluated.append([1, len(ip), len(pc.subsets[0])])
            tr.candidates_scored += 1
            tr.evals += len(pc.subsets[0])
            if octree is not None:
                lv_n[levels[ci] - 1] += 1
                lv_s[levels[ci] - 1] += len(ip)
        if octree is not None:
            # levelscore[level] += E(sc) (fitting.jl:184), summed in closed form per level:
            # sum of E = -n + (N+2)/(M+2) * (sum of sigma + n)
            M1, Nn = len(pc.subsets[0]), pc.size
            for l in range(octree.nlevels):
                if lv_n[l]:
                    levelscore[l] += (-float(lv_n[l])) + (float(Nn + 2) / float(M1 + 2)) * float(lv_s[l] + lv_n[l])
        cc[2] = k * minsubsetN
        tr.sets_drawn = cc[2]
        cc[0] = len(shapes)
        if len(shapes) >= 1:
            if progressive:
                refine_progressive(pc, params, shapes, scores, evaluated, tr)
            best, _ = findhighestscore(scores)
            scr = scores[best].E
            s = cc[sidx[it["extract_s"]]]
            if prob(scr, s, pc.size, drawN) > prob_det:
                if lsq:
                    shapes[best] = lsq_refine(shapes[best], pc, params)[0]
                ip = refit(shapes[best], pc, params)
                if bitmap is not None and len(ip):
                    ip = ip[bitmap_filter(shapes[best], pc.vertices[ip], bitmap[0], bitmap[1])[0]]
                tr.evals += int(pc.isenabled.sum())
                pc.isenabled[ip] = False
                extracted.append(Extracted(shapes[best], ip))
                tr.extracted_at.append(k)
                del shapes[best], scores[best], inpts[best], evaluated[best]
                keep = [j for j in range(len(shapes)) if pc.isenabled[inpts[j]].all()]
                shapes = [shapes[j] for j in keep]
                scores = [scores[j] for j in keep]
                inpts = [inpts[j] for j in keep]
                evaluated = [evaluated[j] for j in keep]
                if progressive:
                    # the survivors' subset-1 inliers are all still enabled, but their inliers in the subsets 2..r may
                    # just have been extracted: a refined score falls back to the (exact) subset-1 score
                    M1 = len(pc.subsets[0])
                    for j in range(len(shapes)):
                        if evaluated[j][0] > 1:
                            evaluated[j] = [1, len(inpts[j]), M1]
                            scores[j] = estimatescore_f64(M1, pc.size, len(inpts[j]))
        if octree is not None:
            if k % max(1, lw_period) == 0:  # lw_period = 1: after every iteration, the reference's schedule
                levelweight = updatelevelweight(levelweight, levelscore)  # iterations.jl:148
            tr.levelweight = levelweight.copy()
        s = cc[sidx[it["terminate_s"]]]
        if prob(tau, s, pc.size, drawN) > prob_det:
            break
    return extracted
