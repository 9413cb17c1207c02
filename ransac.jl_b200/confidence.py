"""ConfidenceInterval and score estimation: mirror of src/confidenceintervals.jl.

`estimatescore` goes through the C ABI (rsc_estimate_score) so that the Int64 wrap-around of the
reference (Q9) is reproduced by the shipped library, not re-derived in Python."""
from __future__ import annotations

import ctypes as C

from . import _lib


class ConfidenceInterval:
    """confidenceintervals.jl:1-6: (min, max, E=(min+max)/2); raises if min > max."""

    __slots__ = ("min", "max", "E")

    def __init__(self, x, y):
        if x > y:
            raise ValueError("out of order")
        self.min = float(x)
        self.max = float(y)
        self.E = (self.min + self.max) / 2

    def __repr__(self):
        return f"CI: [{self.min}, {self.max}]"


def notsoconfident(x, y) -> ConfidenceInterval:
    """confidenceintervals.jl:20-22."""
    if x != x or y != y:
        ci = ConfidenceInterval.__new__(ConfidenceInterval)
        ci.min = ci.max = ci.E = float("nan")
        return ci
    return ConfidenceInterval(min(x, y), max(x, y))


def isoverlap(i1: ConfidenceInterval, i2: ConfidenceInterval) -> bool:
    """confidenceintervals.jl:29-36."""
    if i1.min == i2.min:
        return True
    if i1.min < i2.min:
        return i2.min <= i1.max
    return isoverlap(i2, i1)


def E(x: ConfidenceInterval) -> float:
    """confidenceintervals.jl:43."""
    return x.E


def estimatescore(S1length: int, Plength: int, sigmaS1: int) -> ConfidenceInterval:
    """confidenceintervals.jl:71-74."""
    lo, hi, e = C.c_double(), C.c_double(), C.c_double()
    _lib.lib.rsc_estimate_score(int(S1length), int(Plength), int(sigmaS1), C.byref(lo), C.byref(hi), C.byref(e))
    ci = ConfidenceInterval.__new__(ConfidenceInterval)
    ci.min, ci.max, ci.E = lo.value, hi.value, e.value
    return ci


def estimatescore_f64(Slength: int, Plength: int, sigma: int) -> ConfidenceInterval:
    """The interval confidenceintervals.jl:53-74 means: products in float64, left to right, without the
    reference's Int64 wrap-around (Q9).  Progressive scoring decides on min/max, which the
    wrap corrupts for N >~ 1e6; the reference's own loop only reads E, which it does not."""
    import math

    Np, x, n = float(-2 - Slength), float(-2 - Plength), float(-1 - sigma)
    sq_ = ((x * n) * (Np - x)) * (Np - n) / (Np - 1.0)
    sq = 0.0 if sq_ < 0 else math.sqrt(sq_)
    return notsoconfident(-1 - (x * n + sq) / Np, -1 - (x * n - sq) / Np)


def prob(n, s, N, k):
    """utilities.jl:262."""
    return 1 - (1 - (n / N) ** k) ** s
