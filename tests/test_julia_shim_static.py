"""bindings/julia/KrylovKitB200.jl cannot run here (no Julia in the image), so it is checked STATICALLY against the
header it binds: every `ccall((:name, lib), Ret, (T...), args...)` must name a function `include/b200krylov.h`
declares, with the same number of parameters, a Julia type of the right class for each C parameter, the right
return type, and as many call arguments as parameter types.  This catches what a first run under Julia would —
a missing / swapped / mistyped argument — short of the semantics."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200krylov.h")
SHIM = os.path.join(ROOT, "bindings", "julia", "KrylovKitB200.jl")

# C parameter type (normalised) -> acceptable Julia ccall types
PTR_VOID = {"Ptr{Cvoid}"}
OK = {
    "b2k_ctx*": PTR_VOID, "const b2k_ctx*": PTR_VOID, "b2k_op*": PTR_VOID, "const b2k_op*": PTR_VOID,
    "b2k_ctx**": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"}, "b2k_op**": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "void*": {"Ptr{Cvoid}", "Ptr{T}", "Ptr{Float64}", "Ptr{Float32}"},
    "const void*": {"Ptr{Cvoid}", "Ptr{T}", "Ptr{Float64}", "Ptr{Float32}", "Ptr{Int32}", "Ptr{Int64}", "Ptr{UInt8}", "Ptr{Ti}"},
    "void**": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "int32_t": {"Cint", "Int32"}, "int64_t": {"Int64", "Clonglong"}, "uint64_t": {"UInt64"}, "size_t": {"Csize_t", "UInt64"},
    "double": {"Float64", "Cdouble"},
    "b2k_vec": {"Int32", "Cint"},
    "const b2k_vec*": {"Ptr{Int32}", "Ref{Int32}", "Ptr{Cint}"}, "b2k_vec*": {"Ptr{Int32}", "Ref{Int32}", "Ref{Cint}", "Ptr{Cint}"},
    "int32_t*": {"Ptr{Int32}", "Ref{Int32}", "Ref{Cint}", "Ptr{Cint}"}, "const int32_t*": {"Ptr{Int32}", "Ptr{Cint}"},
    "int64_t*": {"Ptr{Int64}", "Ref{Int64}"}, "const int64_t*": {"Ptr{Int64}"},
    "double*": {"Ptr{Float64}", "Ref{Float64}", "Ref{Cdouble}", "Ptr{Cdouble}"},
    "const double*": {"Ptr{Float64}", "Ref{Float64}", "Ptr{Cdouble}"},
}
RET = {"int32_t": {"Cint", "Int32"}, "const char*": {"Cstring", "Ptr{UInt8}"}, "void*": {"Ptr{Cvoid}"}}


def _strip_comments(txt):
    return re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)


def header_prototypes():
    txt = _strip_comments(open(HEADER).read())
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(b2k_\w+)\s*\(([^;{]*?)\)\s*;", txt):
        ret = " ".join(m.group(1).split()).replace(" *", "*")
        params = []
        body = m.group(3).strip()
        if body and body != "void":
            for p in body.split(","):
                p = " ".join(p.split())
                mm = re.match(r"(.*?)(\b\w+)?$", p)               # drop the parameter name
                typ = mm.group(1).strip() if mm.group(2) and mm.group(1).strip() else p
                typ = typ.replace(" *", "*").replace("* ", "*")
                params.append(typ)
        protos[m.group(2)] = (ret, params)
    return protos


def _split_top(s):
    """split a comma list at nesting depth 0 (parentheses, braces, brackets)"""
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


def shim_ccalls():
    txt = re.sub(r"#.*", "", open(SHIM).read())
    calls = []
    for m in re.finditer(r"ccall\(", txt):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(txt[i], 0)
            i += 1
        parts = _split_top(txt[m.end():i - 1])
        name = re.match(r"\(\s*:(\w+)\s*,\s*lib\s*\)", parts[0]).group(1)
        types = _split_top(parts[2].strip()[1:-1]) if parts[2].strip() != "()" else []
        calls.append((name, parts[1].strip(), [t for t in types if t], parts[3:]))
    return calls


def test_header_parser_sees_the_whole_abi():
    import krylovkit_jl_b200 as kk
    protos = header_prototypes()
    assert set(kk._lib.EXPORTED) <= set(protos)
    assert protos["b2k_op_apply_normal_gram"] == ("int32_t", ["b2k_ctx*", "const b2k_op*", "b2k_vec", "b2k_vec", "b2k_vec"])


def test_every_ccall_matches_its_prototype():
    protos = header_prototypes()
    calls = shim_ccalls()
    assert len(calls) >= 35
    problems = []
    for name, ret, types, args in calls:
        if name not in protos:
            problems.append(f"{name}: not declared in include/b200krylov.h")
            continue
        cret, cparams = protos[name]
        if ret not in RET.get(cret, set()):
            problems.append(f"{name}: returns {cret}, ccall says {ret}")
        if len(types) != len(cparams):
            problems.append(f"{name}: {len(cparams)} C parameters, {len(types)} ccall types")
            continue
        if len(args) != len(types):
            problems.append(f"{name}: {len(types)} ccall types but {len(args)} arguments")
        for k, (ct, jt) in enumerate(zip(cparams, types)):
            if ct not in OK:
                problems.append(f"{name}: parameter {k} has C type '{ct}' this test does not know")
            elif jt not in OK[ct]:
                problems.append(f"{name}: parameter {k} is '{ct}', ccall passes {jt}")
    assert not problems, "\n".join(problems)


def test_the_benchmarked_entry_points_are_bound():
    """the fused path bench.py times and the flagged modes must be reachable from the shim (round-1 verdict)"""
    bound = {c[0] for c in shim_ccalls()}
    for name in ("b2k_lanczos_expand_many", "b2k_op_apply_block", "b2k_block_inner", "b2k_block_reorthogonalize", "b2k_block_qr", "b2k_cg_step", "b2k_cg_chain", "b2k_bicgstab_chain",
                 "b2k_basis_project", "b2k_basis_unproject", "b2k_basis_orthogonalize", "b2k_basis_transform",
                 "b2k_op_apply", "b2k_op_apply_adjoint", "b2k_op_apply_normal_gram"):
        assert name in bound, name
