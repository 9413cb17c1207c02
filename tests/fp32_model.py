"""NumPy float32 model of the tiled kernel's per-pair arithmetic (csrc/rsc_eval.cuh), used on CPU to
calibrate/guard the FP32 error bound `kappa` against the float64 margin.  Test infrastructure."""
import math

import numpy as np

f32 = np.float32
U = 2.0 ** -24
KAPPA = {0: 8.0, 1: 16.0, 2: 16.0, 3: 20.0}  # must match rsc::kappa()


def fma(a, b, c):
    return (a.astype(np.float64) * np.float64(b) + np.asarray(c, np.float64)).astype(f32)


def rsqrt(x):
    # MUFU.RSQ: ~2^-22.4 relative error; documented max rel. error 2^-22.9 = 2.1 u; modelled as a +-2.2 u wiggle
    r = (1.0 / np.sqrt(x.astype(np.float64)))
    return (r * (1 + 2.2 * U * np.sign(np.sin(np.arange(len(x)) * 12.9898)))).astype(f32)


def record(kind, outw, p, pmax, nmax, eps=0.3, cosa=math.cos(math.radians(5))):
    """(fields, band, scale) like rsc::compile_record; the kernel's margin is scale x the plain margin"""
    nm = max(1.0, nmax)
    sg = 1.0 if outw else -1.0
    scale = 1.0
    if kind == 0:
        m = np.array(p[3:6]); mn = np.linalg.norm(m); o = m / mn; oq = float(o @ p[0:3])
        r = [*o, -oq, *(-m)]
        L = (pmax + abs(oq) + 1) * max(mn, 1.0) * nm
    elif kind == 1:
        r = [sg, *(-sg * np.array(p[0:3])), -p[3], -cosa * p[3]]
        L = (pmax + np.linalg.norm(p[0:3]) + abs(p[3]) + 1) * nm
    elif kind == 2:
        r = [sg, *(-sg * np.array(p[3:6])), *p[0:3], -p[6], -cosa * p[6]]
        a2 = float(np.dot(p[0:3], p[0:3]))
        L = (pmax + np.linalg.norm(p[3:6]) + abs(p[6]) + 1) * max(a2, 1.0) * nm
    else:
        ax = np.array(p[3:6]) / np.linalg.norm(p[3:6])  # the reference's cone test only uses the axis direction
        ch, sh = math.cos(p[6] / 2), math.sin(p[6] / 2)
        apex = np.array(p[0:3])
        if ch >= 0.5:  # column type RSC_CONE: margin in units of 1/cos(opang/2)
            scale = 1.0 / ch
            r = [sg, *(-sg * apex), *ax, sg * sh * scale, -eps * scale, cosa * scale]
        else:  # column type kConeWide: margin in units of 1/sin(opang/2)
            assert sh >= 1 / 16, "needle cones are decided in FP64 (infinite band)"
            scale = 1.0 / sh
            r = [sg, *(-sg * apex), *ax, sg * ch * scale, -eps * scale, cosa * scale, ch * scale]
        L = (pmax + np.linalg.norm(apex) + 1) * nm * scale
    return [f32(x) for x in r], KAPPA[kind] * U * L, scale


def margin32(kind, r, P, N, eps, cosa):
    px, py, pz = (P[:, i].astype(f32) for i in range(3))
    nx, ny, nz = (N[:, i].astype(f32) for i in range(3))
    eps, cosa = f32(eps), f32(cosa)
    if kind == 0:
        d = fma(px, r[0], fma(py, r[1], fma(pz, r[2], r[3])))
        e = np.abs(d) - eps
        nt = fma(nx, r[4], fma(ny, r[5], fma(nz, r[6], cosa)))
        return np.maximum(e, nt)
    vx, vy, vz = fma(px, r[0], r[1]), fma(py, r[0], r[2]), fma(pz, r[0], r[3])
    if kind == 1:
        vv = fma(vx, vx, fma(vy, vy, vz * vz))
        d = fma(vv, rsqrt(vv), r[4])
        e = np.abs(d) - eps
        s = fma(vx, nx, fma(vy, ny, fma(vz, nz, r[5])))
        return np.maximum(e, fma(d, cosa, -s))
    h = fma(vx, r[4], fma(vy, r[5], vz * r[6]))
    wx, wy, wz = fma(h, -r[4], vx), fma(h, -r[5], vy), fma(h, -r[6], vz)
    ww = fma(wx, wx, fma(wy, wy, wz * wz))
    if kind == 2:
        d = fma(ww, rsqrt(ww), r[7])
        e = np.abs(d) - eps
        wn = fma(wx, nx, fma(wy, ny, fma(wz, nz, r[8])))
        return np.maximum(e, fma(d, cosa, -wn))
    rho = ww * rsqrt(ww)
    wn = fma(wx, nx, fma(wy, ny, wz * nz))
    if len(r) == 11:  # wide cone
        d = fma(-rho, r[7], h)
        e = np.abs(d) + r[8]
        an = fma(nx, r[4], fma(ny, r[5], nz * r[6]))
        t1 = fma(an, r[0], r[9])
        cw = (r[10] * wn).astype(f32)
        return np.maximum(e, fma(rho, t1, -cw))
    d = fma(h, r[7], -rho)
    e = np.abs(d) + r[8]
    an = fma(nx, r[4], fma(ny, r[5], nz * r[6]))
    t1 = fma(an, r[7], r[9])
    return np.maximum(e, fma(rho, t1, -wn))


def margin64(kind, outw, p, P, N, eps, cosa):
    """float64 margin of the closed forms (equal to the reference's tests up to ~1e-15 relative)."""
    P = P.astype(np.float64); N = N.astype(np.float64)
    sg = 1.0 if outw else -1.0
    p = np.asarray(p, np.float64)
    if kind == 0:
        m = p[3:6]; o = m / np.linalg.norm(m)
        d = (P - p[0:3]) @ o
        return np.maximum(np.abs(d) - eps, cosa - N @ m)
    if kind == 1:
        v = P - p[0:3]; r = np.linalg.norm(v, axis=1)
        return np.maximum(np.abs(r - p[3]) - eps, cosa * r - sg * np.einsum("ij,ij->i", v, N))
    if kind == 2:
        a, c, R = p[0:3], p[3:6], p[6]
        v = P - c; h = v @ a; w = v - h[:, None] * a; rho = np.linalg.norm(w, axis=1)
        return np.maximum(np.abs(rho - R) - eps, cosa * rho - sg * np.einsum("ij,ij->i", w, N))
    apex, a, op = p[0:3], p[3:6] / np.linalg.norm(p[3:6]), p[6]
    v = P - apex; h = v @ a; w = v - h[:, None] * a; rho = np.linalg.norm(w, axis=1)
    c, s = math.cos(op / 2), math.sin(op / 2)
    d = h * s - rho * c
    q = c * np.einsum("ij,ij->i", w, N) - s * rho * (N @ a)
    return np.maximum(np.abs(d) - eps, cosa * rho - sg * q)
