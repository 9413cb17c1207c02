/*
 * oracle.c -- CPU oracle in plain C: scalar float64 restatement of RANSAC.jl's hot path, op for op.
 *
 * TEST INFRASTRUCTURE ONLY (checker for tests/, smoke() and bench.py's cpu_baseline / --impl
 * reference legs).  Nothing in ransac.jl_b200/ links or loads this file.
 *
 * Follows (paths relative to /root/reference, RANSAC.jl v0.6.0):
 *   compatiblesPlane    src/shapes/plane.jl:114-130  (project2plane :82-103)
 *   compatiblesSphere   src/shapes/sphere.jl:144-172
 *   compatiblesCylinder src/shapes/cylinder.jl:194-221
 *   compatiblesCone     src/shapes/cone.jl:132-153   (project2cone :68-85, rodrigues
 *                                                    src/utilities.jl:19-43, :61-64)
 *   fit x4              plane.jl:33-57, sphere.jl:29-114, cylinder.jl:34-168, cone.jl:39-128
 *   estimatescore       src/confidenceintervals.jl:53-74
 * The same operation order as oracle/ransac_oracle.py (tests assert bit-identical masks between
 * the two).  Build with -ffp-contract=off so that no multiply-add is fused.
 *
 * Parity status: see the header of ransac_oracle.py -- PARITY UNPINNED for the compatibles functions, refit and
 * the cylinder/cone fits (no reference test pins them; Julia is not installed here); sphere/plane
 * fit accept-reject answers are pinned by test/dummyspheretest.jl.  LinearAlgebra.rank is restated
 * with a one-sided Jacobi SVD and `\` with partial-pivot Gaussian elimination.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int32_t type;
  int32_t outwards;
  double p[7];
} orc_cand; /* same layout as rsc_cand */

typedef struct {
  double eps[4], alpha[4];
  double parallelthrdeg, sphere_par, minconeopang, collin_threshold;
  int32_t shape_types[4];
  int32_t n_shape_types;
  int32_t sphere_ignores_enabled; /* Q4 */
} orc_params;

typedef struct {
  double x, y, z;
} v3;

static inline v3 vsub(v3 a, v3 b) { return (v3){a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline v3 vadd(v3 a, v3 b) { return (v3){a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline v3 vneg(v3 a) { return (v3){-a.x, -a.y, -a.z}; }
static inline v3 vscale(double s, v3 a) { return (v3){s * a.x, s * a.y, s * a.z}; }
static inline double vdot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline double vnorm(v3 a) { return sqrt(vdot(a, a)); }
static inline v3 vnormalize(v3 a) { return vscale(1.0 / vnorm(a), a); } /* inv(norm(a))*a */
static inline v3 vcross(v3 a, v3 b) {
  return (v3){a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

/* ---- compatibles* ------------------------------------------------------------------------- */
static int compat_plane(const orc_cand* c, v3 p, v3 n, double eps, double thr) {
  v3 q = {c->p[0], c->p[1], c->p[2]}, m = {c->p[3], c->p[4], c->p[5]};
  v3 oz = vnormalize(m);
  double pz = vdot(oz, vsub(p, q));
  return (vdot(m, n) > thr) && (fabs(pz) < eps);
}

static int compat_sphere(const orc_cand* c, v3 p, v3 n, double eps, double thr) {
  v3 o = {c->p[0], c->p[1], c->p[2]};
  double R = c->p[3];
  v3 u = c->outwards ? vnormalize(vsub(p, o)) : vnormalize(vsub(o, p));
  return (vdot(u, n) > thr) && (fabs(vnorm(vsub(p, o)) - R) < eps);
}

static int compat_cylinder(const orc_cand* c, v3 p, v3 n, double eps, double thr) {
  v3 a = {c->p[0], c->p[1], c->p[2]}, ce = {c->p[3], c->p[4], c->p[5]};
  double R = c->p[6];
  double h = vdot(a, vsub(p, ce));
  v3 cn = vsub(vsub(p, vscale(h, a)), ce);
  int okr = fabs(vnorm(cn) - R) < eps;
  v3 u = vnormalize(cn);
  if (!c->outwards) u = vneg(u);
  return okr && (vdot(u, n) > thr);
}

/* project2cone: returns dist, writes the surface normal */
static double project2cone(const orc_cand* c, double ct, double st, v3 p, v3* nrm) {
  v3 apex = {c->p[0], c->p[1], c->p[2]}, axis = {c->p[3], c->p[4], c->p[5]};
  v3 tp = vsub(apex, p);
  v3 tpn = vnormalize(tp);
  v3 rot = vnormalize(vcross(axis, tpn));
  v3 cn = vnormalize(vcross(axis, rot));
  v3 nv = vnormalize(rot);
  double v[3] = {nv.x, nv.y, nv.z};
  double R[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      double o = v[i] * v[j];
      double e = (i == j) ? 1.0 : 0.0;
      R[i][j] = o + ct * (e - o);
    }
  R[0][1] -= st * v[2];
  R[0][2] += st * v[1];
  R[1][0] += st * v[2];
  R[1][2] -= st * v[0];
  R[2][0] -= st * v[1];
  R[2][1] += st * v[0];
  v3 rv = {R[0][0] * cn.x + R[0][1] * cn.y + R[0][2] * cn.z, R[1][0] * cn.x + R[1][1] * cn.y + R[1][2] * cn.z,
           R[2][0] * cn.x + R[2][1] * cn.y + R[2][2] * cn.z};
  v3 cur = vnormalize(rv);
  *nrm = cur;
  return vdot(vneg(cur), vneg(tp));
}

static int compat_cone(const orc_cand* c, double ct, double st, v3 p, v3 n, double eps, double thr) {
  v3 cur;
  double dist = project2cone(c, ct, st, p, &cur);
  v3 nr = c->outwards ? cur : vneg(cur);
  return (vdot(nr, n) > thr) && (fabs(dist) < eps);
}

static int compat_any(const orc_cand* c, double ct, double st, v3 p, v3 n, const double* eps, const double* thr) {
  switch (c->type) {
    case 0:
      return compat_plane(c, p, n, eps[0], thr[0]);
    case 1:
      return compat_sphere(c, p, n, eps[1], thr[1]);
    case 2:
      return compat_cylinder(c, p, n, eps[2], thr[2]);
    case 3:
      return compat_cone(c, ct, st, p, n, eps[3], thr[3]);
  }
  return 0;
}

/*
 * scorecandidate / refit core: for C candidates over n points (AoS float64), count compatible
 * points; `enabled` (nullable, one byte per point) is ANDed in except for spheres when
 * sphere_ignores_enabled (Q4).  masks (nullable): C x n bytes.  Returns the thread count used.
 */
int orc_score(const orc_cand* cands, int C, const double* P, const double* N, int64_t n, const uint8_t* enabled,
              const orc_params* prm, int32_t* counts, uint8_t* masks, int nthreads) {
  double thr[4];
  for (int t = 0; t < 4; ++t) thr[t] = cos(prm->alpha[t]);
  int used = 1;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
  used = nthreads > 0 ? nthreads : omp_get_max_threads();
#endif
  for (int c = 0; c < C; ++c) {
    const orc_cand* cd = &cands[c];
    const double ct = cos(-cd->p[6] / 2), st = sin(-cd->p[6] / 2);
    const int honour = !(cd->type == 1 && prm->sphere_ignores_enabled);
    int64_t cnt = 0;
#pragma omp parallel for reduction(+ : cnt) schedule(static)
    for (int64_t i = 0; i < n; ++i) {
      v3 p = {P[3 * i], P[3 * i + 1], P[3 * i + 2]};
      v3 nn = {N[3 * i], N[3 * i + 1], N[3 * i + 2]};
      int ok = compat_any(cd, ct, st, p, nn, prm->eps, thr);
      if (enabled && honour) ok = ok && enabled[i];
      if (masks) masks[(size_t)c * n + i] = (uint8_t)ok;
      cnt += ok;
    }
    counts[c] = (int32_t)cnt;
  }
  return used;
}

/* ---- fits --------------------------------------------------------------------------------- */
static int fit_plane(const v3* p, const v3* n, int k, const orc_params* prm, orc_cand* out) {
  v3 cr = vnormalize(vcross(vsub(p[1], p[0]), vsub(p[2], p[0])));
  if (vnorm(cr) < prm->collin_threshold) return 0; /* Q3: never fires for finite input */
  double thr = cos(prm->alpha[0]);
  int all_ok = 1, all_inv = 1;
  for (int i = 0; i < k; ++i) {
    double d = vdot(cr, vnormalize(n[i]));
    all_ok = all_ok && (d > thr);
    all_inv = all_inv && (d < -thr);
  }
  if (!all_ok && !all_inv) return 0;
  if (!all_ok) cr = vscale(-1.0, cr);
  out->type = 0;
  out->outwards = 1;
  out->p[0] = p[0].x, out->p[1] = p[0].y, out->p[2] = p[0].z;
  out->p[3] = cr.x, out->p[4] = cr.y, out->p[5] = cr.z, out->p[6] = 0.0;
  return 1;
}

static double cosd_(double deg) { return cos(deg * (M_PI / 180.0)); }

static int fit_sphere(const v3* v, const v3* n, int k, const orc_params* prm, orc_cand* out) {
  v3 n1n = vnormalize(n[0]), n2n = vnormalize(n[1]);
  v3 ctr;
  double R;
  if (fabs(vdot(n1n, n2n)) > cosd_(prm->parallelthrdeg)) {
    ctr = vscale(0.5, vadd(v[0], v[1]));
    /* (v1+v2)/2 : division by 2 == multiplication by 0.5 exactly */
    R = vnorm(vsub(ctr, v[0]));
  } else {
    v3 g = vsub(v[1], v[0]);
    v3 h = vcross(n2n, g), kk = vcross(n2n, n1n);
    double nk = vnorm(kk), nh = vnorm(h);
    if (nk < prm->sphere_par || nh < prm->sphere_par) {
      v3 n2 = vcross(n2n, vcross(n1n, n2n));
      v3 n1 = vcross(n1n, vcross(n2n, n1n));
      v3 c1 = vadd(v[0], vscale(vdot(vsub(v[1], v[0]), n2) / vdot(n[0], n2), n[0]));
      v3 c2 = vadd(v[1], vscale(vdot(vsub(v[0], v[1]), n1) / vdot(n[1], n1), n[1]));
      ctr = vscale(0.5, vadd(c1, c2));
      R = (vnorm(vsub(v[0], ctr)) + vnorm(vsub(v[0], ctr))) / 2; /* Q5 */
    } else {
      double f = nh / nk;
      ctr = (vdot(h, kk) > 0) ? vadd(v[0], vscale(f, n1n)) : vsub(v[0], vscale(f, n1n));
      R = vnorm(vsub(ctr, v[0]));
    }
  }
  double thr = cos(prm->alpha[1]);
  int vert = 1, ok = 1, inv = 1;
  for (int i = 0; i < k; ++i) {
    vert = vert && (fabs(vnorm(vsub(v[i], ctr)) - R) < prm->eps[1]);
    double d = vdot(vnormalize(vsub(v[i], ctr)), vnormalize(n[i]));
    ok = ok && (d > thr);
    inv = inv && (d < -thr);
  }
  if (!vert || (!ok && !inv)) return 0;
  out->type = 1;
  out->outwards = ok ? 1 : 0;
  out->p[0] = ctr.x, out->p[1] = ctr.y, out->p[2] = ctr.z, out->p[3] = R;
  out->p[4] = out->p[5] = out->p[6] = 0.0;
  return 1;
}

static v3 proj2plane(v3 n, v3 w) { return vadd(w, vscale(vdot(vneg(n), w) / vdot(n, n), n)); }

static void projectto2d(v3 xa, v3 ya, v3 za, v3 p, double* r) {
  double xx = xa.x, xy = xa.y, xz = xa.z, yx = ya.x, yy = ya.y, yz = ya.z, zx = za.x, zy = za.y, zz = za.z;
  double px = p.x, py = p.y, pz = p.z;
  double den = xz * yy * zx - xy * yz * zx - xz * yx * zy + xx * yz * zy + xy * yx * zz - xx * yy * zz;
  double n1 = -(pz * yy * zx) + py * yz * zx + pz * yx * zy - px * yz * zy - py * yx * zz + px * yy * zz;
  double n2 = pz * xy * zx - py * xz * zx - pz * xx * zy + px * xz * zy + py * xx * zz - px * xy * zz;
  r[0] = -(n1 / den);
  r[1] = -(n2 / den);
}

static int fit_cylinder(const v3* p, const v3* n, int k, const orc_params* prm, orc_cand* out) {
  if (fabs(vdot(n[0], n[1])) > cosd_(prm->parallelthrdeg)) return 0; /* Q8 */
  v3 an = vnormalize(vcross(n[0], n[1]));
  v3 xax = vnormalize(proj2plane(an, p[0]));
  v3 yax = vnormalize(vcross(an, xax));
  double a[2], b[2], c[2], d[2];
  projectto2d(xax, yax, an, proj2plane(an, p[0]), a);
  projectto2d(xax, yax, an, proj2plane(an, vadd(p[0], n[0])), b);
  projectto2d(xax, yax, an, proj2plane(an, p[1]), c);
  projectto2d(xax, yax, an, proj2plane(an, vadd(p[1], n[1])), d);
  double amb[2] = {a[0] - b[0], a[1] - b[1]}, cmd[2] = {c[0] - d[0], c[1] - d[1]};
  double d1 = a[0] * b[1] - a[1] * b[0];
  double d2 = c[0] * d[1] - c[1] * d[0];
  double d3 = amb[0] * cmd[1] - amb[1] * cmd[0];
  double ic0 = (d1 * cmd[0] - d2 * amb[0]) / d3, ic1 = (d1 * cmd[1] - d2 * amb[1]) / d3;
  v3 ce = vadd(vscale(ic0, xax), vscale(ic1, yax));
  double nn[2];
  for (int i = 0; i < 2; ++i) {
    v3 pc = vsub(p[i], ce);
    nn[i] = vnorm(vsub(pc, vscale(vdot(an, pc), an)));
  }
  double R = (nn[0] + nn[1]) / 2;
  double thr = cos(prm->alpha[2]);
  int vert = 1, ok = 1, inv = 1;
  for (int i = 0; i < k; ++i) {
    v3 cn = vsub(vsub(p[i], vscale(vdot(an, vsub(p[i], ce)), an)), ce);
    vert = vert && (fabs(vnorm(cn) - R) < prm->eps[2]);
    double dd = vdot(vnormalize(cn), n[i]);
    ok = ok && (dd > thr);
    inv = inv && (dd < -thr);
  }
  if (!vert || (!ok && !inv)) return 0;
  out->type = 2;
  out->outwards = ok ? 1 : 0;
  out->p[0] = an.x, out->p[1] = an.y, out->p[2] = an.z;
  out->p[3] = ce.x, out->p[4] = ce.y, out->p[5] = ce.z, out->p[6] = R;
  return 1;
}

/* singular values of an r x c matrix (r = 3, c = 3 or 4) by one-sided Jacobi on the transpose */
static void singular_values(const double* A, int r, int c, double* s) {
  /* work on B = A (r rows); orthogonalise the ROWS pairwise: singular values = row norms */
  double B[3][4];
  for (int i = 0; i < r; ++i)
    for (int j = 0; j < c; ++j) B[i][j] = A[i * c + j];
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < r - 1; ++i)
      for (int j = i + 1; j < r; ++j) {
        double aii = 0, ajj = 0, aij = 0;
        for (int t = 0; t < c; ++t) aii += B[i][t] * B[i][t], ajj += B[j][t] * B[j][t], aij += B[i][t] * B[j][t];
        if (aij == 0.0) continue;
        off = fmax(off, fabs(aij) / sqrt(aii * ajj + 1e-300));
        double zeta = (ajj - aii) / (2.0 * aij);
        double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        double cs = 1.0 / sqrt(1.0 + tt * tt), sn = cs * tt;
        for (int t = 0; t < c; ++t) {
          double bi = B[i][t], bj = B[j][t];
          B[i][t] = cs * bi - sn * bj;
          B[j][t] = sn * bi + cs * bj;
        }
      }
    if (off < 1e-17) break;
  }
  for (int i = 0; i < r; ++i) {
    double q = 0;
    for (int t = 0; t < c; ++t) q += B[i][t] * B[i][t];
    s[i] = sqrt(q);
  }
}

static int rank_julia(const double* A, int r, int c) {
  for (int i = 0; i < r * c; ++i)
    if (!isfinite(A[i])) return -1;
  double s[3];
  singular_values(A, r, c, s);
  double smax = fmax(s[0], fmax(s[1], s[2]));
  double tol = (r < c ? r : c) * 2.220446049250313e-16 * smax;
  int rk = 0;
  for (int i = 0; i < r; ++i) rk += s[i] > tol;
  return rk;
}

/* x = A \ b, 3x3, Gaussian elimination with partial pivoting */
static int solve3(const double* A, const double* b, double* x) {
  double M[3][4];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) M[i][j] = A[i * 3 + j];
    M[i][3] = b[i];
  }
  for (int col = 0; col < 3; ++col) {
    int piv = col;
    for (int r = col + 1; r < 3; ++r)
      if (fabs(M[r][col]) > fabs(M[piv][col])) piv = r;
    if (M[piv][col] == 0.0) return 0;
    if (piv != col)
      for (int j = 0; j < 4; ++j) {
        double t = M[col][j];
        M[col][j] = M[piv][j];
        M[piv][j] = t;
      }
    for (int r = col + 1; r < 3; ++r) {
      double f = M[r][col] / M[col][col];
      for (int j = col; j < 4; ++j) M[r][j] -= f * M[col][j];
    }
  }
  for (int i = 2; i >= 0; --i) {
    double s = M[i][3];
    for (int j = i + 1; j < 3; ++j) s -= M[i][j] * x[j];
    x[i] = s / M[i][i];
  }
  return 1;
}

static double clamp1(double x) { return x != x ? x : (x < -1 ? -1 : (x > 1 ? 1 : x)); }

static int fit_cone(const v3* p, const v3* n, int k, const orc_params* prm, orc_cand* out) {
  double r[9] = {n[0].x, n[0].y, n[0].z, n[1].x, n[1].y, n[1].z, n[2].x, n[2].y, n[2].z};
  if (rank_julia(r, 3, 3) != 3) return 0;
  double ds[3] = {vdot(p[0], n[0]), vdot(p[1], n[1]), vdot(p[2], n[2])};
  double rv[12];
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) rv[i * 4 + j] = r[i * 3 + j];
    rv[i * 4 + 3] = -1 * ds[i];
  }
  if (rank_julia(rv, 3, 4) != 3) return 0;
  double apx[3];
  if (!solve3(r, ds, apx)) return 0;
  v3 ap = {apx[0], apx[1], apx[2]};
  v3 a3[3];
  for (int i = 0; i < 3; ++i) {
    v3 d = vsub(p[i], ap);
    double nd = vnorm(d);
    a3[i] = vadd(ap, (v3){d.x / nd, d.y / nd, d.z / nd});
  }
  v3 ax = vnormalize(vcross(vsub(a3[1], a3[0]), vsub(a3[2], a3[0])));
  v3 sm = vadd(vadd(a3[0], a3[1]), a3[2]);
  v3 midp = {sm.x / 3, sm.y / 3, sm.z / 3};
  v3 dirv = vnormalize(vsub(midp, ap));
  if (vdot(ax, dirv) < 0) ax = vscale(-1.0, ax);
  double ang[3];
  for (int i = 0; i < 3; ++i) ang[i] = acos(clamp1(vdot(vnormalize(vsub(p[i], ap)), ax)));
  double op = 2 * (ang[0] + ang[1] + ang[2]) / 3;
  orc_cand c;
  c.type = 3;
  c.outwards = 1;
  c.p[0] = ap.x, c.p[1] = ap.y, c.p[2] = ap.z, c.p[3] = ax.x, c.p[4] = ax.y, c.p[5] = ax.z, c.p[6] = op;
  /* validatecone */
  double ct = cos(-op / 2), st = sin(-op / 2);
  v3 nr[8];
  if (k > 8) k = 8;
  for (int i = 0; i < k; ++i)
    if (project2cone(&c, ct, st, p[i], &nr[i]) > prm->eps[3]) return 0; /* Q6: signed */
  if (op < prm->minconeopang) return 0;
  double thr = cos(prm->alpha[3]);
  int ok = 1, inv = 1;
  for (int i = 0; i < k; ++i) {
    double d = vdot(nr[i], n[i]);
    ok = ok && (d > thr);
    inv = inv && (d < -thr);
  }
  if (!ok && !inv) return 0;
  c.outwards = ok ? 1 : 0;
  *out = c;
  return 1;
}

/* forcefitshapes! for ONE minimal set (fitting.jl:165-173): candidates in shape_types order; returns how many */
static int fit_set(const v3* p, const v3* n, int k, const orc_params* prm, orc_cand* out) {
  int m = 0;
  for (int t = 0; t < prm->n_shape_types; ++t) {
    orc_cand c;
    memset(&c, 0, sizeof(c));
    int ok = 0;
    switch (prm->shape_types[t]) {
      case 0:
        ok = fit_plane(p, n, k, prm, &c);
        break;
      case 1:
        ok = fit_sphere(p, n, k, prm, &c);
        break;
      case 2:
        ok = fit_cylinder(p, n, k, prm, &c);
        break;
      case 3:
        ok = fit_cone(p, n, k, prm, &c);
        break;
    }
    if (ok) out[m++] = c;
  }
  return m;
}

/*
 * forcefitshapes! for S minimal sets: P, N = S x k x 3 doubles.  Candidates are written compacted in
 * (set, shape_types) order; out_set[i] = source set.  Returns the number of candidates.
 */
int orc_fit_points(const double* P, const double* N, int S, int k, const orc_params* prm, orc_cand* out, int32_t* out_set) {
  int m = 0;
  if (k > 8) return -1;
  for (int s = 0; s < S; ++s) {
    v3 p[8], n[8];
    for (int i = 0; i < k; ++i) {
      const double* pp = P + ((size_t)s * k + i) * 3;
      const double* nn = N + ((size_t)s * k + i) * 3;
      p[i] = (v3){pp[0], pp[1], pp[2]};
      n[i] = (v3){nn[0], nn[1], nn[2]};
    }
    orc_cand c[4];
    int got = fit_set(p, n, k, prm, c);
    for (int j = 0; j < got; ++j) {
      out[m] = c[j];
      out_set[m] = s;
      ++m;
    }
  }
  return m;
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* =====================================================================================
 * The loop: ransac(pc, params) -- src/iterations.jl:35-162, with
 *   samplepointcloud4!      src/fitting.jl:383-430 (root cell only: Q1, octree.jl:82-84)
 *   forcefitshapes!         src/fitting.jl:165-173
 *   scorecandidates!        src/fitting.jl:181-190 (subset 1 only: iterations.jl:95)
 *   scorecandidate x4       plane.jl:61-71, sphere.jl:118-135 (Q4: ignores isenabled),
 *                           cylinder.jl:172-186, cone.jl:155-167
 *   findhighestscore        src/fitting.jl:140-150 (strict >, first maximum wins)
 *   estimatescore           src/confidenceintervals.jl:53-74 (Int64 products wrap: Q9)
 *   prob / chooseS          src/utilities.jl:262, :297-300
 *   refit x4                plane.jl:137-143, sphere.jl:179-190, cylinder.jl:228-234, cone.jl:176-182
 *   invalidate_indexes!     src/fitting.jl:197-202
 *   removeinvalidshapes!    src/fitting.jl:209-221 (stored inlier lists, literally)
 * Julia's random stream is version dependent (Q19) and not part of the contract: the minimal set
 * `(k-1)*minsubsetN + i` draws from a Philox4x32-10 stream keyed by the seed, exactly like
 * oracle/ransac_oracle.py::SetStream / sample_minimal_set (tests assert C loop == NumPy loop).
 * OpenMP only parallelises over independent sets / (candidate, point) pairs; every result is
 * assembled in the reference's sequential order.
 * ===================================================================================== */

typedef struct {
  int32_t drawN, minsubsetN, itermax, extract_s, terminate_s, reserved;
  int64_t tau;
  double prob_det;
} orc_iter;

typedef struct orc_run {
  int nshapes, cap;
  orc_cand* shapes;
  int64_t* len;
  int64_t** idx;
  int32_t* extracted_at;
  int iterations;
  int64_t cands_scored, evals;
  double seconds_sample_fit, seconds_score, seconds_extract;
} orc_run;

/* ---- Philox4x32-10 set streams (ransac_oracle.py:927-960) ---- */
static void philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0, c1 = n1, c2 = n2, c3 = n3;
    k0 += 0x9E3779B9u, k1 += 0xBB67AE85u;
  }
  out[0] = c0, out[1] = c1, out[2] = c2, out[3] = c3;
}

typedef struct {
  uint32_t key[2];
  uint64_t set_id;
  uint32_t ndraw;
  uint32_t blk[4];
} set_stream;

static uint64_t stream_u64(set_stream* s) {
  const uint32_t i = s->ndraw++;
  if ((i & 1u) == 0) {
    uint32_t ctr[4] = {i / 2, (uint32_t)s->set_id, (uint32_t)(s->set_id >> 32), 0};
    philox4x32(ctr, s->key, s->blk);
    return (uint64_t)s->blk[0] | ((uint64_t)s->blk[1] << 32);
  }
  return (uint64_t)s->blk[2] | ((uint64_t)s->blk[3] << 32);
}

static int64_t rand_below(set_stream* s, int64_t n) { return (int64_t)(((unsigned __int128)stream_u64(s) * (unsigned __int128)(uint64_t)n) >> 64); }

/* samplepointcloud4! on the root cell; returns 1 and idx[drawN] on success */
static int sample_set(int64_t n, const uint8_t* enabled, const int64_t* en_idx, int64_t ne, int drawN, uint64_t seed, uint64_t set_id,
                      int64_t* idx) {
  set_stream st = {{(uint32_t)seed, (uint32_t)(seed >> 32)}, set_id, 0, {0, 0, 0, 0}};
  if (ne == 0) return 0; /* the reference would spin forever; the tau test keeps it from getting here */
  int64_t r1 = rand_below(&st, n);
  while (!enabled[r1]) r1 = rand_below(&st, n);
  if (ne < drawN) return 0;
  idx[0] = r1;
  for (int k = 1; k < drawN; ++k) {
    int64_t nexti = rand_below(&st, ne);
    if (idx[0] == en_idx[nexti]) nexti = rand_below(&st, ne); /* "try oncemore", fitting.jl:417-420 */
    idx[k] = en_idx[nexti];
  }
  for (int i = 1; i < drawN; ++i) /* allisdifferent, utilities.jl:285-295 */
    for (int j = 0; j < i; ++j)
      if (idx[i] == idx[j]) return 0;
  return 1;
}

/* estimatescore with Julia's Int64 arithmetic (products wrap, Q9); E = (min+max)/2 */
static int64_t wrap_mul(int64_t a, int64_t b) { return (int64_t)((uint64_t)a * (uint64_t)b); }
void orc_estimate_score(int64_t s1len, int64_t plen, int64_t sigma, double* omin, double* omax, double* oE) {
  const int64_t N = -2 - s1len, x = -2 - plen, n = -1 - sigma;
  const int64_t xn = wrap_mul(x, n);
  const int64_t prod = wrap_mul(wrap_mul(xn, N - x), N - n);
  const double sq_ = (double)prod / (double)(N - 1);
  const double sq = sq_ < 0 ? 0.0 : sqrt(sq_);
  const double gmin = ((double)xn + sq) / (double)N, gmax = ((double)xn - sq) / (double)N;
  const double a = -1 - gmin, b = -1 - gmax;
  const double lo = a < b ? a : b, hi = a < b ? b : a; /* notsoconfident */
  if (omin) *omin = lo;
  if (omax) *omax = hi;
  if (oE) *oE = (lo + hi) / 2;
}

static double prob_(double n, double s, double N, double k) { return 1 - pow(1 - pow(n / N, k), s); }

static double now_s(void) {
#ifdef _OPENMP
  return omp_get_wtime();
#else
  return 0.0;
#endif
}

typedef struct {
  orc_cand c;
  double E;
  int32_t* in; /* inpoints: global indices in subset order */
  int64_t nin;
} stored;

/*
 * P, N: n x 3 float64; subset1: m global indices (pc.subsets[1]); enabled: n bytes, in/out
 * (ransac(pc, params, setenabled): the caller sets them to 1 for setenabled = true).
 */
orc_run* orc_ransac(const double* P, const double* Nrm, int64_t n, const int64_t* subset1, int64_t m, uint8_t* enabled,
                    const orc_params* prm, const orc_iter* it, uint64_t seed, int nthreads) {
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
  orc_run* run = (orc_run*)calloc(1, sizeof(orc_run));
  const int drawN = it->drawN, S = it->minsubsetN, nt = prm->n_shape_types;
  if (drawN < 2 || drawN > 8 || n >= 2147483647LL) return run;
  double thr[4];
  for (int t = 0; t < 4; ++t) thr[t] = cos(prm->alpha[t]);
  int64_t n_enabled = 0;
  for (int64_t i = 0; i < n; ++i) n_enabled += enabled[i] != 0;
  int64_t* en_idx = (int64_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
  int64_t ne = 0;
  int en_dirty = 1;
  stored* store = NULL;
  int64_t nstore = 0, capstore = 0;
  orc_cand* newc = (orc_cand*)malloc((size_t)S * nt * sizeof(orc_cand));
  int32_t* newn = (int32_t*)malloc((size_t)S * sizeof(int32_t));
  const int CB = 256; /* candidates scored per block (bounds the byte-mask scratch) */
  uint8_t* mk = (uint8_t*)malloc((size_t)CB * (size_t)(m > 0 ? m : 1));
  uint8_t* rmask = (uint8_t*)malloc((size_t)(n > 0 ? n : 1));
  int64_t cc[3] = {0, 0, 0}; /* lengthC, allcand, nofminset: iterations.jl:70 */

  for (int k = 1; k <= it->itermax; ++k) {
    if (n_enabled < it->tau) break; /* iterations.jl:75 */
    run->iterations = k;
    double t0 = now_s();
    if (en_dirty) {
      ne = 0;
      for (int64_t i = 0; i < n; ++i)
        if (enabled[i]) en_idx[ne++] = i;
      en_dirty = 0;
    }
    /* ---- minsubsetN minimal sets -> candidates in (set, shape_types) order ---- */
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < S; ++i) {
      int64_t sd[8];
      newn[i] = 0;
      if (!sample_set(n, enabled, en_idx, ne, drawN, seed, (uint64_t)(k - 1) * (uint64_t)S + (uint64_t)i, sd)) continue;
      v3 p[8], nn[8];
      for (int j = 0; j < drawN; ++j) {
        p[j] = (v3){P[3 * sd[j]], P[3 * sd[j] + 1], P[3 * sd[j] + 2]};
        nn[j] = (v3){Nrm[3 * sd[j]], Nrm[3 * sd[j] + 1], Nrm[3 * sd[j] + 2]};
      }
      newn[i] = fit_set(p, nn, drawN, prm, newc + (size_t)i * nt);
    }
    int64_t ncand = 0;
    for (int i = 0; i < S; ++i) {
      for (int j = 0; j < newn[i]; ++j) newc[ncand + j] = newc[(size_t)i * nt + j]; /* ncand <= i*nt: in-place compaction */
      ncand += newn[i];
    }
    cc[1] += ncand; /* iterations.jl:94 */
    double t1 = now_s();
    run->seconds_sample_fit += t1 - t0;
    /* ---- scorecandidates! on subset 1 ---- */
    if (nstore + ncand > capstore) {
      capstore = (nstore + ncand) * 2 + 64;
      store = (stored*)realloc(store, (size_t)capstore * sizeof(stored));
    }
    for (int64_t c0 = 0; c0 < ncand; c0 += CB) {
      const int nb = (int)(ncand - c0 < CB ? ncand - c0 : CB);
      const int64_t chunk = 8192, nch = (m + chunk - 1) / chunk;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
      for (int c = 0; c < nb; ++c)
        for (int64_t ch = 0; ch < nch; ++ch) {
          const orc_cand* cd = &newc[c0 + c];
          const double ct = cos(-cd->p[6] / 2), st = sin(-cd->p[6] / 2);
          const int honour = !(cd->type == 1 && prm->sphere_ignores_enabled);
          const int64_t hi = (ch + 1) * chunk < m ? (ch + 1) * chunk : m;
          uint8_t* row = mk + (size_t)c * (size_t)m;
          for (int64_t j = ch * chunk; j < hi; ++j) {
            const int64_t g = subset1[j];
            v3 p = {P[3 * g], P[3 * g + 1], P[3 * g + 2]}, nn = {Nrm[3 * g], Nrm[3 * g + 1], Nrm[3 * g + 2]};
            int ok = compat_any(cd, ct, st, p, nn, prm->eps, thr);
            if (honour) ok = ok && enabled[g];
            row[j] = (uint8_t)ok;
          }
        }
#pragma omp parallel for schedule(dynamic, 1)
      for (int c = 0; c < nb; ++c) {
        const uint8_t* row = mk + (size_t)c * (size_t)m;
        int64_t cnt = 0;
        for (int64_t j = 0; j < m; ++j) cnt += row[j];
        stored* sp = &store[nstore + c];
        sp->c = newc[c0 + c];
        sp->nin = cnt;
        sp->in = (int32_t*)malloc((size_t)(cnt > 0 ? cnt : 1) * sizeof(int32_t));
        int64_t o = 0;
        for (int64_t j = 0; j < m; ++j)
          if (row[j]) sp->in[o++] = (int32_t)subset1[j];
        orc_estimate_score(m, n, cnt, NULL, NULL, &sp->E);
      }
      nstore += nb;
    }
    run->cands_scored += ncand;
    run->evals += ncand * m;
    cc[2] = (int64_t)k * S; /* iterations.jl:99 */
    cc[0] = nstore;         /* iterations.jl:102 */
    double t2 = now_s();
    run->seconds_score += t2 - t1;
    if (nstore >= 1) {
      int64_t best = 0; /* findhighestscore: strict >, first wins */
      double highest = store[0].E;
      for (int64_t i = 0; i < nstore; ++i)
        if (store[i].E > highest) highest = store[i].E, best = i;
      const double s_ex = (double)cc[it->extract_s];
      if (prob_(highest, s_ex, (double)n, (double)drawN) > it->prob_det) {
        /* ---- refit over the enabled points of the whole cloud, invalidate_indexes! ---- */
        const orc_cand bc = store[best].c;
        const double ct = cos(-bc.p[6] / 2), st = sin(-bc.p[6] / 2);
        int64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
        for (int64_t i = 0; i < n; ++i) {
          int ok = 0;
          if (enabled[i]) {
            v3 p = {P[3 * i], P[3 * i + 1], P[3 * i + 2]}, nn = {Nrm[3 * i], Nrm[3 * i + 1], Nrm[3 * i + 2]};
            ok = compat_any(&bc, ct, st, p, nn, prm->eps, thr);
          }
          rmask[i] = (uint8_t)ok;
          total += ok;
        }
        if (run->nshapes == run->cap) {
          run->cap = run->cap ? 2 * run->cap : 16;
          run->shapes = (orc_cand*)realloc(run->shapes, (size_t)run->cap * sizeof(orc_cand));
          run->len = (int64_t*)realloc(run->len, (size_t)run->cap * sizeof(int64_t));
          run->idx = (int64_t**)realloc(run->idx, (size_t)run->cap * sizeof(int64_t*));
          run->extracted_at = (int32_t*)realloc(run->extracted_at, (size_t)run->cap * sizeof(int32_t));
        }
        int64_t* list = (int64_t*)malloc((size_t)(total > 0 ? total : 1) * sizeof(int64_t));
        int64_t o = 0;
        for (int64_t i = 0; i < n; ++i)
          if (rmask[i]) list[o++] = i, enabled[i] = 0;
        n_enabled -= total;
        en_dirty = 1;
        run->shapes[run->nshapes] = bc, run->len[run->nshapes] = total, run->idx[run->nshapes] = list;
        run->extracted_at[run->nshapes] = k;
        ++run->nshapes;
        run->evals += ne;
        /* deleteat!(scoredshapes, best) + removeinvalidshapes! */
        uint8_t* dead = (uint8_t*)calloc((size_t)nstore, 1);
        dead[best] = 1;
#pragma omp parallel for schedule(dynamic, 16)
        for (int64_t i = 0; i < nstore; ++i) {
          if (i == best) continue;
          const stored* sp = &store[i];
          for (int64_t j = 0; j < sp->nin; ++j)
            if (!enabled[sp->in[j]]) {
              dead[i] = 1;
              break;
            }
        }
        int64_t w = 0;
        for (int64_t i = 0; i < nstore; ++i) {
          if (dead[i])
            free(store[i].in);
          else
            store[w++] = store[i];
        }
        nstore = w;
        free(dead);
      }
    }
    run->seconds_extract += now_s() - t2;
    /* iterations.jl:151-156 */
    if (prob_((double)it->tau, (double)cc[it->terminate_s], (double)n, (double)drawN) > it->prob_det) break;
  }
  for (int64_t i = 0; i < nstore; ++i) free(store[i].in);
  free(store), free(newc), free(newn), free(mk), free(rmask), free(en_idx);
  return run;
}

int orc_run_nshapes(const orc_run* r) { return r->nshapes; }
int orc_run_iterations(const orc_run* r) { return r->iterations; }
int64_t orc_run_cands_scored(const orc_run* r) { return r->cands_scored; }
int64_t orc_run_evals(const orc_run* r) { return r->evals; }
void orc_run_seconds(const orc_run* r, double* out3) {
  out3[0] = r->seconds_sample_fit, out3[1] = r->seconds_score, out3[2] = r->seconds_extract;
}
int64_t orc_run_shape(const orc_run* r, int i, orc_cand* out, int32_t* extracted_at) {
  if (i < 0 || i >= r->nshapes) return -1;
  if (out) *out = r->shapes[i];
  if (extracted_at) *extracted_at = r->extracted_at[i];
  return r->len[i];
}
void orc_run_inpoints(const orc_run* r, int i, int64_t* out) {
  if (i < 0 || i >= r->nshapes) return;
  memcpy(out, r->idx[i], (size_t)r->len[i] * sizeof(int64_t));
}
void orc_run_free(orc_run* r) {
  if (!r) return;
  for (int i = 0; i < r->nshapes; ++i) free(r->idx[i]);
  free(r->shapes), free(r->len), free(r->idx), free(r->extracted_at), free(r);
}
