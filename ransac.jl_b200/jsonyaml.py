"""JSON export of shapes and YAML import of parameters: host-side mirror of src/json-yaml.jl.

`toDict` / `exportJSON` (json-yaml.jl:9-56) and `readconfig` / `dict2nt` (json-yaml.jl:78-110) with
the reference's key names: a shape's dict has "type" (= strt) plus one entry per struct field, a
vector of shapes becomes {"primitives": [...]}; config files use the reference's Greek keys
(ϵ, α, τ), which map to this package's `eps`, `alpha`, `tau`.  Pinned by the reference's own tests
(test/json.jl, test/yaml.jl + test/yaml/t1.yml, t2.yml) in tests/test_jsonyaml.py.
"""
from __future__ import annotations

import dataclasses
import json
from typing import Any, Dict, Sequence, Union

import numpy as np
import yaml

from .params import DEFAULT_PARAMETERS, DEFAULT_SHAPE_DICT, ransacparameters
from .shapes import ExtractedShape, FittedShape, strt

_GREEK = {"ϵ": "eps", "ε": "eps", "α": "alpha", "τ": "tau"}


def _plain(v: Any):
    if isinstance(v, np.ndarray):
        return [float(x) for x in v]
    if isinstance(v, (np.floating,)):
        return float(v)
    if isinstance(v, (np.bool_,)):
        return bool(v)
    if isinstance(v, (np.integer,)):
        return int(v)
    return v


def toDict(s: Union[FittedShape, ExtractedShape, Sequence]) -> Dict[str, Any]:
    """json-yaml.jl:9-31: {"type": strt(s), <every field of the struct>}; vectors -> {"primitives": [...]}"""
    if isinstance(s, ExtractedShape):
        return toDict(s.shape)
    if isinstance(s, FittedShape):
        d: Dict[str, Any] = {"type": strt(s)}
        for f in dataclasses.fields(s):
            d[f.name] = _plain(getattr(s, f.name))
        return d
    return {"primitives": [toDict(x) for x in s]}


def exportJSON(io, s, indent=None) -> None:
    """json-yaml.jl:44-56: print a shape / extracted shape / vector of them to `io` as JSON."""
    json.dump(toDict(s), io, indent=indent, separators=None if indent else (",", ":"), ensure_ascii=False)
    if indent:
        io.write("\n")


def dict2nt(entries) -> dict:
    """json-yaml.jl:100-110: a YAML list of one-key maps -> one merged group"""
    out: dict = {}
    for d in entries:
        for k, v in d.items():
            out[_GREEK.get(k, k)] = v
    return out


def readconfig(fname, toextend=None, shapedict=None) -> dict:
    """json-yaml.jl:78-98: read a YAML config over a base parameter set (default DEFAULT_PARAMETERS)."""
    p = DEFAULT_PARAMETERS if toextend is None else toextend
    shapedict = DEFAULT_SHAPE_DICT if shapedict is None else shapedict
    with open(fname, "r", encoding="utf-8") as f:
        fdict = yaml.safe_load(f)
    for group, entries in fdict.items():
        nt = dict2nt(entries)
        if "shape_types" in nt:
            nt["shape_types"] = [shapedict[k] for k in nt["shape_types"]]
        p = ransacparameters(p, **{group: nt})
    return p
