"""Static lint of the Julia binding (julia/RANSACB200/src/RANSACB200.jl) and of the `ccall` snippets in
INTEGRATION.md against include/rsc.h.  The image has no Julia, so the binding is never executed here; this
keeps every `ccall` it contains consistent with the header: the symbol exists, the argument-type tuple has
the header's arity, each Julia type is one the C parameter can be bound with, the return type matches, and
the call passes as many values as it declares types.  The POD mirrors are checked field by field against
the ctypes structs the GPU tests use."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINDING = os.path.join(ROOT, "julia", "RANSACB200", "src", "RANSACB200.jl")


def _header_prototypes():
    """name -> (return type, [parameter types]) from include/rsc.h (comments stripped)"""
    src = open(os.path.join(ROOT, "include", "rsc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = {}
    for m in re.finditer(r"^\s*([A-Za-z_][\w \*]*?)\s*\b(rsc_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", src, flags=re.M | re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        if ret.startswith("typedef"):
            continue
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        ptypes = []
        for a in params:
            a = re.sub(r"\bconst\b", "", a).strip()
            t = re.sub(r"\s*\b\w+$", "", a) if re.search(r"[\w\*]\s+\**\w+$|\*\w+$", a) else a  # drop the parameter name
            ptypes.append(t.replace(" ", ""))
        protos[name] = (ret.replace("const ", "").replace(" ", ""), ptypes)
    return protos


# which Julia ccall types may bind a C parameter type
OK = {
    "int32_t": {"Int32", "Cint"},
    "int64_t": {"Int64"},
    "uint64_t": {"UInt64"},
    "double": {"Float64", "Cdouble"},
    "void*": {"Ptr{Cvoid}", "Ptr{UInt8}"},
    "rsc_ctx*": {"Ptr{Cvoid}"},
    "rsc_cloud*": {"Ptr{Cvoid}"},
    "rsc_run*": {"Ptr{Cvoid}"},
    "rsc_ctx**": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "rsc_cloud**": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "rsc_run**": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "rsc_params*": {"Ref{RscParams}", "Ptr{RscParams}"},
    "rsc_cand*": {"Ref{RscCand}", "Ptr{RscCand}"},
    "float*": {"Ptr{Float32}"},
    "double*": {"Ptr{Float64}", "Ref{Float64}"},
    "int32_t*": {"Ptr{Int32}", "Ref{Int32}"},
    "int64_t*": {"Ptr{Int64}", "Ref{Int64}"},
    "uint32_t*": {"Ptr{UInt32}"},
    "uint64_t*": {"Ptr{UInt64}"},
}
RET = {"int32_t": {"Int32", "Cint"}, "int64_t": {"Int64"}, "double": {"Float64", "Cdouble"}, "void": {"Cvoid"}, "char*": {"Cstring"}}


def _split_top(s):
    """split at top-level commas (parentheses / braces / brackets nest)"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _ccalls(text):
    """(symbol, return type, [argument types], number of values passed) of every ccall in a Julia source text"""
    calls = []
    for m in re.finditer(r"ccall\(", text):
        i, depth = m.end(), 1
        while depth:  # the matching parenthesis of this ccall
            depth += {"(": 1, ")": -1}.get(text[i], 0)
            i += 1
        parts = _split_top(" ".join(text[m.end(): i - 1].split()))
        sym = re.search(r":(rsc_[a-z0-9_]+)", parts[0])
        if not sym:
            continue
        tup = parts[2].strip()
        assert tup.startswith("(") and tup.endswith(")"), (sym.group(1), tup)
        types = [t for t in _split_top(tup[1:-1]) if t]
        calls.append((sym.group(1), parts[1], types, len(parts) - 3))
    return calls


def _check(calls, protos, where):
    assert calls, f"no ccall found in {where}"
    for name, ret, types, nvals in calls:
        assert name in protos, f"{where}: {name} is not declared in include/rsc.h"
        cret, cargs = protos[name]
        assert ret in RET[cret], f"{where}: {name} returns {cret}, bound as {ret}"
        assert len(types) == len(cargs), f"{where}: {name} takes {len(cargs)} arguments, ccall declares {len(types)}"
        assert nvals == len(types), f"{where}: {name}: {len(types)} argument types but {nvals} values"
        for k, (jt, ct) in enumerate(zip(types, cargs)):
            assert jt in OK[ct], f"{where}: {name} argument {k + 1} is `{ct}`, bound as `{jt}`"


def test_every_ccall_of_the_binding_matches_the_header():
    protos = _header_prototypes()
    assert len(protos) >= 50
    calls = _ccalls(open(BINDING).read())
    assert len(calls) >= 25
    _check(calls, protos, "RANSACB200.jl")
    # the entry points a drop-in needs are all bound
    bound = {c[0] for c in calls}
    for need in ("rsc_ctx_create", "rsc_cloud_create", "rsc_cloud_create_f64", "rsc_cloud_set_subset", "rsc_cloud_set_enabled",
                 "rsc_cloud_get_enabled", "rsc_score", "rsc_fit_batch", "rsc_refit_extract", "rsc_ransac_run", "rsc_run_shape",
                 "rsc_run_inpoints", "rsc_run_destroy", "rsc_comm_unique_id", "rsc_ctx_comm_init", "rsc_last_error"):
        assert need in bound, f"{need} is not bound by the Julia package"


def test_ccall_snippets_of_integration_md_match_the_header():
    protos = _header_prototypes()
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```julia\n(.*?)```", text, flags=re.S)
    calls = [c for b in blocks for c in _ccalls(b)]
    assert len(calls) >= 8
    _check(calls, protos, "INTEGRATION.md")


def _julia_struct(name):
    src = open(BINDING).read()
    body = re.search(r"struct %s\b(.*?)\nend" % name, src, flags=re.S).group(1)
    fields = []
    for line in body.splitlines():
        line = line.split("#")[0].strip()
        m = re.match(r"(\w+)::(.+)$", line)
        if m:
            fields.append((m.group(1), m.group(2).strip()))
    return fields


def _ctype_of(jt):
    base = {"Int32": C.c_int32, "UInt32": C.c_uint32, "Int64": C.c_int64, "Float64": C.c_double}
    m = re.match(r"NTuple\{(\d+),\s*(\w+)\}", jt)
    return base[m.group(2)] * int(m.group(1)) if m else base[jt]


@pytest.mark.parametrize("jname,cname", [("RscCand", "rsc_cand"), ("RscParams", "rsc_params")])
def test_pod_mirrors_of_the_binding_match_the_ctypes_structs(jname, cname):
    import ransac_jl_b200 as R

    cstruct = getattr(R._lib, cname)
    jf = _julia_struct(jname)
    cf = list(cstruct._fields_)
    assert [f for f, _ in jf] == [f for f, _ in cf], "field names / order differ from the ctypes mirror of include/rsc.h"

    class J(C.Structure):  # the layout Julia gives an isbits struct = the C layout of the same field types
        _fields_ = [(f, _ctype_of(t)) for f, t in jf]

    assert C.sizeof(J) == C.sizeof(cstruct)
    for f, _ in jf:
        assert getattr(J, f).offset == getattr(cstruct, f).offset and getattr(J, f).size == getattr(cstruct, f).size, f


def test_loader_does_not_use_a_non_constant_library_tuple():
    """`ccall((:sym, lib), ...)` needs a compile-time constant `lib` on older Julia 1.x; the binding passes dlsym pointers"""
    src = open(BINDING).read()
    assert not re.search(r"ccall\(\s*\(\s*:rsc_", src)
    assert "Libdl.dlsym" in src and "Libdl.dlopen" in src
