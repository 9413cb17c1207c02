"""Fitted-shape types: mirror of FittedPlane/FittedSphere/FittedCylinder/FittedCone
(src/shapes/plane.jl:8-11, sphere.jl:9-13, cylinder.jl:11-16, cone.jl:11-19) and ExtractedShape
(src/fitting.jl:81-84), with the conversion to/from the 64-byte C-ABI candidate record."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import _lib


class FittedShape:
    """Abstract base (fitting.jl:8): user shapes subclass this and plug into the host loop."""

    kind = -1

    def to_cand(self) -> _lib.rsc_cand:  # pragma: no cover - abstract
        raise NotImplementedError


def _v3(x):
    a = np.asarray(x, dtype=np.float64).reshape(3)
    return a


@dataclass
class FittedPlane(FittedShape):
    point: np.ndarray
    normal: np.ndarray
    kind = _lib.RSC_PLANE

    def __post_init__(self):
        self.point, self.normal = _v3(self.point), _v3(self.normal)

    def to_cand(self):
        c = _lib.rsc_cand(type=self.kind, outwards=1)
        c.p[:] = [*self.point, *self.normal, 0.0]
        return c


@dataclass
class FittedSphere(FittedShape):
    center: np.ndarray
    radius: float
    outwards: bool
    kind = _lib.RSC_SPHERE

    def __post_init__(self):
        self.center = _v3(self.center)

    def to_cand(self):
        c = _lib.rsc_cand(type=self.kind, outwards=int(bool(self.outwards)))
        c.p[:] = [*self.center, float(self.radius), 0.0, 0.0, 0.0]
        return c


@dataclass
class FittedCylinder(FittedShape):
    axis: np.ndarray
    center: np.ndarray
    radius: float
    outwards: bool
    kind = _lib.RSC_CYLINDER

    def __post_init__(self):
        self.axis, self.center = _v3(self.axis), _v3(self.center)

    def to_cand(self):
        c = _lib.rsc_cand(type=self.kind, outwards=int(bool(self.outwards)))
        c.p[:] = [*self.axis, *self.center, float(self.radius)]
        return c


@dataclass
class FittedCone(FittedShape):
    apex: np.ndarray
    axis: np.ndarray
    opang: float
    outwards: bool
    kind = _lib.RSC_CONE

    def __post_init__(self):
        self.apex, self.axis = _v3(self.apex), _v3(self.axis)

    def to_cand(self):
        c = _lib.rsc_cand(type=self.kind, outwards=int(bool(self.outwards)))
        c.p[:] = [*self.apex, *self.axis, float(self.opang)]
        return c


SHAPE_KIND = {FittedPlane: 0, FittedSphere: 1, FittedCylinder: 2, FittedCone: 3}
KIND_SHAPE = {v: k for k, v in SHAPE_KIND.items()}


def strt(s: FittedShape) -> str:
    """strt (plane.jl:19, sphere.jl:22, cylinder.jl:24, cone.jl:27)."""
    return {0: "plane", 1: "sphere", 2: "cylinder", 3: "cone"}[s.kind]


def from_cand(c: _lib.rsc_cand) -> FittedShape:
    p = list(c.p)
    if c.type == _lib.RSC_PLANE:
        return FittedPlane(p[0:3], p[3:6])
    if c.type == _lib.RSC_SPHERE:
        return FittedSphere(p[0:3], p[3], bool(c.outwards))
    if c.type == _lib.RSC_CYLINDER:
        return FittedCylinder(p[0:3], p[3:6], p[6], bool(c.outwards))
    if c.type == _lib.RSC_CONE:
        return FittedCone(p[0:3], p[3:6], p[6], bool(c.outwards))
    raise ValueError(f"unknown shape type {c.type}")


def pack_cands(shapes: Sequence[FittedShape]):
    """ctypes array of rsc_cand for a list of shapes."""
    arr = (_lib.rsc_cand * max(len(shapes), 1))()
    for i, s in enumerate(shapes):
        arr[i] = s.to_cand()
    return arr


def cands_to_numpy(arr, n: int) -> np.ndarray:
    """View n rsc_cand records as a (n, 8) float64 array (column 0 packs type/outwards)."""
    return np.frombuffer(arr, dtype=np.float64, count=8 * n).reshape(n, 8)


@dataclass
class ExtractedShape:
    """fitting.jl:81-84 -- `inpoints` are ascending 0-based global indices here."""

    shape: FittedShape
    inpoints: np.ndarray
