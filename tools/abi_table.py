#!/usr/bin/env python
"""The C ABI, parsed three ways and held equal: the prototypes in include/bn_cuda.h, every `ccall((:bn_*, libbn_cuda),
...)` in julia/src/**, and the ctypes table of bayesnets.jl_b200/_lib.py.

    python tools/abi_table.py            # markdown table for INTEGRATION.md (generated, not hand-written)
    python tools/abi_table.py --check    # exit 1 on any arity / type disagreement

tests/test_host.py::test_julia_ccalls_match_the_header runs the same check on CPU: the Julia glue cannot be executed in
this image, so this is what keeps it from drifting away from the header.
"""
import ctypes as C
import glob
import os
import re
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
HEADER = os.path.join(ROOT, "include", "bn_cuda.h")
JULIA = sorted(glob.glob(os.path.join(ROOT, "julia", "src", "**", "*.jl"), recursive=True))

HANDLES = {"bn_net_t", "bn_factor_t", "bn_comm_t"}
# C scalar -> the Julia type a ccall must use
SCALAR = {"int32_t": "Int32", "uint32_t": "UInt32", "int64_t": "Int64", "uint64_t": "UInt64", "double": "Float64",
          "uint8_t": "UInt8", "int8_t": "Int8", "char": "UInt8", "void": "Cvoid"}
SCALAR.update({h: "UInt64" for h in HANDLES})


def split_top(s, sep=","):
    """Split on `sep` at bracket depth 0."""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == sep and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return [x.strip() for x in out]


def c_type(decl):
    """'const int32_t* card' -> ('int32_t', 1): base type and pointer depth (the parameter name is dropped)."""
    d = decl.replace("const", " ").strip()
    depth = d.count("*")
    d = d.replace("*", " ")
    toks = d.split()
    base = toks[0]
    return base, depth


def header_prototypes():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)\b(bn_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", txt):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        args = [] if params in ("", "void") else [c_type(p) for p in split_top(params)]
        names = [] if params in ("", "void") else [p.replace("*", " ").split()[-1] for p in split_top(params)]
        protos[name] = {"ret": c_type(ret + " x"), "args": args, "names": names, "decl": " ".join(m.group(0).split())}
    return protos


def julia_ok(c, jl):
    """Is the Julia ccall type `jl` a correct binding of the C type c = (base, pointer depth)?"""
    base, depth = c
    jl = jl.replace(" ", "")
    if depth == 0:
        return jl == SCALAR[base]
    if base == "void" and depth == 1:
        return jl in ("Ptr{Cvoid}", "Ptr{Nothing}")
    if base == "void" and depth == 2:
        return jl in ("Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}")
    if base == "char" and depth == 1:
        return jl in ("Cstring", "Ptr{UInt8}", "Ptr{Cchar}")
    inner = SCALAR[base]
    return depth == 1 and jl in (f"Ptr{{{inner}}}", f"Ref{{{inner}}}")


def julia_ccalls():
    calls = []
    for path in JULIA:
        src = open(path).read()
        for m in re.finditer(r"ccall\(\(:(bn_[a-z0-9_]+),\s*libbn_cuda\)", src):
            # balanced scan to the end of this ccall
            i, depth = m.start() + len("ccall"), 0
            j = i
            while True:
                ch = src[j]
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                    if depth == 0:
                        break
                j += 1
            parts = split_top(src[i + 1:j])          # [(:name, lib), Ret, (ArgTypes...), args...]
            tup = parts[2].strip()
            assert tup.startswith("(") and tup.endswith(")"), (path, m.group(1), tup)
            types = split_top(tup[1:-1])
            types = [t for t in types if t]          # a one-element tuple ends with a comma
            line = src.count("\n", 0, m.start()) + 1
            calls.append({"name": m.group(1), "ret": parts[1], "types": types, "n_args": len(parts) - 3,
                          "where": f"{os.path.relpath(path, ROOT)}:{line}"})
    return calls


def ctypes_ok(c, ct):
    base, depth = c
    scal = {"int32_t": C.c_int32, "uint32_t": C.c_uint32, "int64_t": C.c_int64, "uint64_t": C.c_uint64, "double": C.c_double,
            "uint8_t": C.c_uint8, "int8_t": C.c_int8}
    scal.update({h: C.c_uint64 for h in HANDLES})
    if depth == 0:
        return ct is scal[base]
    if ct is C.c_void_p or ct is C.c_char_p:
        return True                                  # an untyped device / host pointer
    if depth == 2:
        return ct in (C.POINTER(C.c_void_p),)
    return base in scal and ct is C.POINTER(scal[base])


def check():
    sys.path.insert(0, ROOT)
    from bayesnets_jl_b200 import _lib
    protos = header_prototypes()
    errs = []
    # header <-> ctypes
    for name, p in protos.items():
        if name not in _lib._SIGS:
            errs.append(f"{name}: declared in bn_cuda.h, missing from _lib._SIGS")
            continue
        res, args = _lib._SIGS[name]
        if len(args) != len(p["args"]):
            errs.append(f"{name}: ctypes binds {len(args)} arguments, header declares {len(p['args'])}")
            continue
        for k, (c, ct) in enumerate(zip(p["args"], args)):
            if not ctypes_ok(c, ct):
                errs.append(f"{name}: ctypes argument {k} ({p['names'][k]}) is {ct}, header says {c}")
    for name in _lib._SIGS:
        if name not in protos:
            errs.append(f"{name}: bound by ctypes, not declared in bn_cuda.h")
    # header <-> Julia
    calls = julia_ccalls()
    for c in calls:
        p = protos.get(c["name"])
        if p is None:
            errs.append(f"{c['where']}: ccall of {c['name']}, which bn_cuda.h does not declare")
            continue
        if len(c["types"]) != len(p["args"]):
            errs.append(f"{c['where']}: {c['name']} ccall lists {len(c['types'])} argument types, header declares {len(p['args'])}")
            continue
        if c["n_args"] != len(c["types"]):
            errs.append(f"{c['where']}: {c['name']} ccall passes {c['n_args']} values for {len(c['types'])} types")
        if not julia_ok(p["ret"], c["ret"]):
            errs.append(f"{c['where']}: {c['name']} return type {c['ret']}, header says {p['ret']}")
        for k, (ct, jt) in enumerate(zip(p["args"], c["types"])):
            if not julia_ok(ct, jt):
                errs.append(f"{c['where']}: {c['name']} argument {k} ({p['names'][k]}) is {jt}, header says {ct}")
    return protos, calls, errs


def markdown(protos, calls):
    by = {}
    for c in calls:
        by.setdefault(c["name"], []).append(c["where"])
    lines = ["| C entry point (include/bn_cuda.h) | arguments | Julia `ccall` sites | ctypes |", "|---|---|---|---|"]
    for name, p in protos.items():
        args = ", ".join(f"{b}{'*' * d} {n}" for (b, d), n in zip(p["args"], p["names"])) or "void"
        sites = "<br>".join(by.get(name, [])) or "-- (host mirror only)"
        lines.append(f"| `{name}` | `{args}` | {sites} | `_lib._SIGS[\"{name}\"]` |")
    return "\n".join(lines)


if __name__ == "__main__":
    protos, calls, errs = check()
    if "--check" in sys.argv:
        for e in errs:
            print("ABI MISMATCH:", e)
        print(f"{len(protos)} prototypes, {len(calls)} ccall sites over {len({c['name'] for c in calls})} entry points, "
              f"{len(errs)} mismatches")
        sys.exit(1 if errs else 0)
    print(markdown(protos, calls))
