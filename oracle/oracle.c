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
      if (ok) {
        out[m] = c;
        out_set[m] = s;
        ++m;
      }
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
