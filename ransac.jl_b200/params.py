"""Parameter system: host-side mirror of RANSAC.jl's nested NamedTuples (src/utilities.jl:302-504).

Parameters are nested dicts with the reference's group names (`iteration`, `common`, `plane`,
`sphere`, `cylinder`, `cone`); `eps`/`alpha`/`tau` spell the reference's ϵ/α/τ.  `to_c()` flattens
them into the `rsc_params` POD that crosses the C ABI.
"""
from __future__ import annotations

import math
from typing import Optional

from . import _lib
from .shapes import FittedCone, FittedCylinder, FittedPlane, FittedSphere, SHAPE_KIND

_S = {"lengthC": _lib.RSC_S_LENGTHC, "allcand": _lib.RSC_S_ALLCAND, "nofminset": _lib.RSC_S_NOFMINSET}


def defaultshapeparameters(shape_type) -> dict:
    """defaultshapeparameters(::Type{S}) -- plane.jl:21, sphere.jl:24, cylinder.jl:26, cone.jl:29."""
    if shape_type is FittedPlane:
        return {"plane": {"eps": 0.3, "alpha": math.radians(5)}}
    if shape_type is FittedSphere:
        return {"sphere": {"eps": 0.3, "alpha": math.radians(5), "sphere_par": 0.02}}
    if shape_type is FittedCylinder:
        return {"cylinder": {"eps": 0.3, "alpha": math.radians(5)}}
    if shape_type is FittedCone:
        return {"cone": {"eps": 0.3, "alpha": math.radians(5), "minconeopang": math.radians(2)}}
    # user-defined shapes bring their own method, like `defaultshapeparameters(::Type{MyShape})` in the
    # reference (docs/src/newprimitive.md:12-18, fitting.jl:15)
    own = getattr(shape_type, "defaultshapeparameters", None)
    if callable(own):
        return dict(own())
    raise TypeError(f"no default parameters for {shape_type!r}: define a static `defaultshapeparameters()` on the class")


def defaultiterationparameters(shape_types) -> dict:
    """utilities.jl:332-347."""
    return {
        "iteration": {
            "drawN": 3,
            "minsubsetN": 15,
            "prob_det": 0.9,
            "shape_types": list(shape_types),
            "tau": 900,
            "itermax": 1000,
            "extract_s": "nofminset",
            "terminate_s": "nofminset",
        }
    }


def defaultcommonparameters() -> dict:
    """utilities.jl:368-373."""
    return {"common": {"collin_threshold": 0.2, "parallelthrdeg": 1.0}}


def defaultparameters(shape_types) -> dict:
    """utilities.jl:391-399."""
    p = defaultiterationparameters(shape_types)
    p.update(defaultcommonparameters())
    for s in shape_types:
        p.update(defaultshapeparameters(s))
    return p


#: RANSAC.jl:94
DEFAULT_SHAPE_TYPES = [FittedPlane, FittedCone, FittedCylinder, FittedSphere]
DEFAULT_PARAMETERS = defaultparameters(DEFAULT_SHAPE_TYPES)
#: RANSAC.jl:100
DEFAULT_SHAPE_DICT = {"plane": FittedPlane, "cone": FittedCone, "cylinder": FittedCylinder, "sphere": FittedSphere}


def ransacparameters(p=None, **kwargs) -> dict:
    """ransacparameters (utilities.jl:425-433, 461-464): override groups of an existing tuple, or
    build the defaults for a list of shape types first."""
    if p is None:
        p = DEFAULT_PARAMETERS
    elif isinstance(p, (list, tuple)):
        p = defaultparameters(p)
    new = {k: dict(v) for k, v in p.items()}
    for k, v in kwargs.items():
        old = dict(p.get(k, v))
        old.update(v)
        new[k] = old
    return new


def setfloattype(nt: dict, T) -> dict:
    """setfloattype(nt, T) (utilities.jl:488-504): convert every real, non-integer number of a nested
    parameter dict to the float type `T` (np.float32 / np.float64 / float); integers, strings and
    everything else pass through.  Used with `RANSACCloud(...; force_eltype=T)`."""
    import numbers

    out = {}
    for k, v in nt.items():
        if isinstance(v, dict):
            out[k] = setfloattype(v, T)
        elif isinstance(v, numbers.Real) and not isinstance(v, (numbers.Integral, bool)):
            out[k] = T(v)
        else:
            out[k] = v
    return out


def to_c(params: dict, compat_flags: Optional[int] = None) -> _lib.rsc_params:
    """Flatten the nested parameters into the C-ABI POD (include/rsc.h, rsc_params)."""
    c = _lib.rsc_params()
    _lib.lib.rsc_params_default(c)
    it = params["iteration"]
    c.drawN = int(it["drawN"])
    c.minsubsetN = int(it["minsubsetN"])
    c.prob_det = float(it["prob_det"])
    c.tau = int(it["tau"])
    c.itermax = int(it["itermax"])
    c.extract_s = _S[it["extract_s"]]
    c.terminate_s = _S[it["terminate_s"]]
    types = it["shape_types"]
    if len(types) > 4:
        raise ValueError("at most the four built-in shape types can be passed to the device loop")
    c.n_shape_types = len(types)
    for i, t in enumerate(types):
        c.shape_types[i] = SHAPE_KIND[t]
    cm = params.get("common", {})
    c.collin_threshold = float(cm.get("collin_threshold", 0.2))
    c.parallelthrdeg = float(cm.get("parallelthrdeg", 1.0))
    for name, kind in (("plane", 0), ("sphere", 1), ("cylinder", 2), ("cone", 3)):
        g = params.get(name)
        if g is not None:
            c.eps[kind] = float(g["eps"])
            c.alpha[kind] = float(g["alpha"])
    if "sphere" in params:
        c.sphere_par = float(params["sphere"].get("sphere_par", 0.02))
    if "cone" in params:
        c.minconeopang = float(params["cone"].get("minconeopang", math.radians(2)))
    if compat_flags is not None:
        c.compat_flags = compat_flags
    return c
