"""The Julia extension cannot run here (no julia in the image), so its `ccall`s are checked statically
against include/makb200.h: every symbol it binds is declared, and the number of argument types in each
ccall tuple equals the number of parameters of the C prototype (and of the ctypes signature the Python
mirror uses, which IS exercised)."""
import os
import re

import makb200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _c_prototypes():
    txt = open(os.path.join(ROOT, "include", "makb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(makb200_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", txt, flags=re.S):
        args = m.group(2).strip()
        protos[m.group(1)] = 0 if args in ("", "void") else len(_split_top(args))
    return protos


def _c_param_classes():
    """name -> list of 'int' | 'size' | 'double' | 'ptr' per parameter."""
    txt = open(os.path.join(ROOT, "include", "makb200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(makb200_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", txt, flags=re.S):
        args = m.group(2).strip()
        cls = []
        for a in ([] if args in ("", "void") else _split_top(args)):
            a = " ".join(a.split())
            if "*" in a or "cudaStream_t" in a:
                cls.append("ptr")
            elif a.startswith("size_t"):
                cls.append("size")
            elif a.startswith("double"):
                cls.append("double")
            elif a.startswith("int") or a.startswith("unsigned"):
                cls.append("int")
            else:
                cls.append("?")
        out[m.group(1)] = cls
    return out


def _julia_class(t):
    t = t.strip()
    if t in ("Cint", "Cuint"):
        return "int"
    if t == "Csize_t":
        return "size"
    if t == "Cdouble":
        return "double"
    if t.startswith(("Ptr{", "CuPtr{", "Ref{")) or t in ("CUDA.CUstream", "Cstring"):
        return "ptr"
    return "?"


def _julia_ccalls():
    calls = []
    for fn in ("yab200.jl", "MatrixAlgebraKitB200Ext.jl"):
        txt = open(os.path.join(ROOT, "ext", "MatrixAlgebraKitB200Ext", fn)).read()
        for m in re.finditer(r"ccall\(\(:(makb200_[a-z0-9_]+),\s*libmakb200\),\s*([A-Za-z_{}.]+),\s*\(", txt):
            # the argument-type tuple starts at m.end()-1; find its matching parenthesis
            i, depth = m.end() - 1, 0
            for j in range(i, len(txt)):
                depth += txt[j] == "("
                depth -= txt[j] == ")"
                if depth == 0:
                    break
            tup = txt[i + 1:j]
            types = [t for t in _split_top(tup) if t]
            # the call's actual arguments: from after the tuple to the ccall's closing parenthesis
            k, depth = j + 1, 1
            for e in range(j + 1, len(txt)):
                depth += txt[e] == "("
                depth -= txt[e] == ")"
                if depth == 0:
                    break
            args = [a for a in _split_top(txt[k:e].lstrip(", \n")) if a]
            calls.append((fn, m.group(1), len(types), len(args), types))
    return calls


def test_julia_ccalls_match_the_header_and_ctypes():
    protos = _c_prototypes()
    calls = _julia_ccalls()
    assert len(calls) >= 35, len(calls)
    classes = _c_param_classes()
    for fn, name, ntypes, nargs, types in calls:
        assert name in protos, f"{fn}: {name} is not declared in makb200.h"
        assert ntypes == protos[name], f"{fn}: ccall of {name} lists {ntypes} types, the C prototype has {protos[name]} parameters"
        assert nargs == ntypes, f"{fn}: ccall of {name} passes {nargs} arguments for {ntypes} types"
        jl = [_julia_class(t) for t in types]
        assert "?" not in jl and "?" not in classes[name], (name, types, classes[name])
        assert jl == classes[name], f"{fn}: ccall of {name}: Julia types {jl} vs C parameters {classes[name]}"
        sig = makb200._lib.SIGNATURES[name]
        assert len(sig[1]) == protos[name], f"ctypes signature of {name} has {len(sig[1])} arguments, the header {protos[name]}"


def _ctypes_class(t):
    import ctypes as C
    if t in (C.c_int, C.c_uint):
        return "int"
    if t is C.c_size_t:
        return "size"
    if t is C.c_double:
        return "double"
    if t is C.c_void_p or t is C.c_char_p or hasattr(t, "_type_") and not isinstance(t._type_, str):
        return "ptr"
    return "?"


def test_every_ctypes_signature_matches_the_header():
    protos = _c_prototypes()
    classes = _c_param_classes()
    for name, (_, args) in makb200._lib.SIGNATURES.items():
        assert name in protos, name
        assert len(args) == protos[name], (name, len(args), protos[name])
        got = [_ctypes_class(t) for t in args]
        assert got == classes[name], f"ctypes signature of {name}: {got} vs C parameters {classes[name]}"
