"""The drop-in boundary seen from its three bindings: include/coldatoms_b200.h compiled as plain C (gcc -std=c11 -pedantic -Werror: no C++,
no torch types), the ctypes structs of the Python mirror (_abi.py) and the `struct`s of the Julia shim (julia/NeutralAtomsB200.jl) must give
every field of ca_beam / ca_model / ca_ode_opts / ca_stats the same offset and every struct the same size.  The probe is a C program generated
from the ctypes field names (a name the header does not declare fails to compile); it also resolves the library's entry points from C and
checks that ca_init refuses to run without a GPU (no CPU fallback behind the C ABI either)."""
import ctypes as C
import json
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STRUCTS = ["ca_beam", "ca_model", "ca_ode_opts", "ca_stats"]
JULIA = {"CaBeam": "ca_beam", "CaModel": "ca_model", "CaOde": "ca_ode_opts", "CaStats": "ca_stats"}


@pytest.fixture(scope="module")
def probe(pkg, tmp_path_factory):
    d = tmp_path_factory.mktemp("abi")
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include <dlfcn.h>', '#include "coldatoms_b200.h"', 'int main(int argc, char** argv) {',
             '  printf("{");']
    for s in STRUCTS:
        st = getattr(pkg._abi, s)
        lines.append('  printf("\\"%s\\": {\\"size\\": %%zu, \\"fields\\": {", sizeof(%s));' % (s, s))
        for i, (name, _) in enumerate(st._fields_):
            lines.append('  printf("%s\\"%s\\": %%zu", offsetof(%s, %s));' % ("" if i == 0 else ", ", name, s, name))
        lines.append('  printf("}}, ");')
    lines += ['  void* h = dlopen(argv[1], RTLD_NOW);',
              '  if (!h) { printf("\\"dlopen\\": \\"%s\\"}", dlerror()); return 0; }',
              '  int (*version)(void); int (*init)(int, ca_ctx**); const char* (*last_error)(void);',
              '  *(void**)(&version) = dlsym(h, "ca_version"); *(void**)(&init) = dlsym(h, "ca_init"); *(void**)(&last_error) = dlsym(h, "ca_last_error");',
              '  (void)sizeof(version == &ca_version); (void)sizeof(init == &ca_init); (void)sizeof(last_error == &ca_last_error);   /* prototypes as declared */',
              '  ca_ctx* ctx = NULL;',
              '  int rc = init(0, &ctx);',
              '  printf("\\"version\\": %d, \\"init_rc\\": %d, \\"ctx_null\\": %d, \\"error_len\\": %zu}", version(), rc, ctx == NULL, strlen(last_error()));',
              '  return 0; }']
    src = d / "abi_probe.c"
    src.write_text('#include <string.h>\n' + "\n".join(lines) + "\n")
    exe = d / "abi_probe"
    subprocess.check_call(["gcc", "-std=c11", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src), "-ldl"])
    out = subprocess.run([str(exe), pkg._lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    return json.loads(out)


def test_header_is_plain_c_and_matches_the_ctypes_mirror(pkg, probe):
    for s in STRUCTS:
        st = getattr(pkg._abi, s)
        assert probe[s]["size"] == C.sizeof(st), s
        for name, _ in st._fields_:
            assert probe[s]["fields"][name] == getattr(st, name).offset, (s, name)


def test_entry_points_resolve_from_c_and_refuse_without_a_gpu(pkg, probe):
    import torch
    assert "dlopen" not in probe, probe.get("dlopen")
    assert probe["version"] == 200
    if torch.cuda.is_available():
        assert probe["init_rc"] == 0 and not probe["ctx_null"]
    else:
        assert probe["init_rc"] == pkg._abi.CA_ERR_NODEVICE and probe["ctx_null"] and probe["error_len"] > 0


def _julia_structs():
    src = open(os.path.join(ROOT, "coldatoms.jl_b200", "julia", "NeutralAtomsB200.jl")).read()
    out = {}
    for m in re.finditer(r"^(?:mutable )?struct (\w+)\n(.*?)^end", src, re.S | re.M):
        if m.group(1) not in JULIA:
            continue
        fields = []
        for line in m.group(2).splitlines():
            line = line.split("#")[0]
            if "=" in line and "::" not in line.split("=")[0]:
                continue   # inner constructor
            for f in line.split(";"):
                fm = re.match(r"\s*(\w+)::([\w{},]+)\s*$", f)
                if fm:
                    fields.append((fm.group(1), fm.group(2)))
        out[m.group(1)] = fields
    return out


def test_julia_shim_structs_have_the_c_layout(probe):
    """C-compatible layout of the Julia `struct`s (natural alignment, nested isbits structs inline) against the C compiler's offsets."""
    js = _julia_structs()
    assert set(js) == set(JULIA)
    prim = {"Float64": (8, 8), "Int64": (8, 8), "UInt64": (8, 8), "Int32": (4, 4), "NTuple{3,Float64}": (24, 8)}

    def layout(name):
        off, align, offs = 0, 1, []
        for fname, ftype in js[name]:
            size, al = prim[ftype] if ftype in prim else layout(ftype)[:2]
            off = (off + al - 1) // al * al
            offs.append((fname, off))
            off += size
            align = max(align, al)
        return (off + align - 1) // align * align, align, offs

    for jname, cname in JULIA.items():
        size, _, offs = layout(jname)
        assert size == probe[cname]["size"], jname
        c_offsets = sorted(probe[cname]["fields"].values())
        assert [o for _, o in offs] == c_offsets, jname
        for fname, o in offs:   # same names wherever the shim keeps the header's name
            if fname in probe[cname]["fields"]:
                assert probe[cname]["fields"][fname] == o, (jname, fname)


def _c_prototypes():
    hdr = open(os.path.join(ROOT, "include", "coldatoms_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    hdr = re.sub(r"//[^\n]*", " ", hdr)
    protos = {}
    for m in re.finditer(r"\b(ca_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", hdr, re.S):
        params = [p.strip() for p in m.group(2).split(",")]
        if params == ["void"] or params == [""]:
            params = []
        kinds = []
        for p in params:
            if "*" in p:
                kinds.append("ptr")
            elif re.search(r"\buint64_t\b|unsigned long long", p):
                kinds.append("u64")
            elif re.search(r"\bint64_t\b|\blong long\b", p):
                kinds.append("i64")
            elif re.search(r"\bdouble\b", p):
                kinds.append("f64")
            elif re.search(r"\bu?int32_t\b|\bint\b|\bunsigned\b", p):
                kinds.append("i32")
            else:
                kinds.append("?" + p)
        protos[m.group(1)] = kinds
    return protos


def _julia_kind(t):
    t = t.strip()
    if t.startswith(("Ptr{", "Ref{")) or t == "Cstring":
        return "ptr"
    return {"Int32": "i32", "Cint": "i32", "UInt32": "i32", "Int64": "i64", "UInt64": "u64", "Float64": "f64"}.get(t, "?" + t)


def _split_top(s):   # split on commas outside braces / parentheses
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def test_julia_ccall_signatures_match_the_header():
    """The shim cannot run here (no Julia), so every `ccall` in it is checked statically: the argument-type tuple must have the arity of the C
    prototype it binds and the same kind (pointer / 32-bit / 64-bit signed / 64-bit unsigned / double) in every position.  Call sites that go
    through `entry(:single, :multi)` are checked against both prototypes."""
    src = open(os.path.join(ROOT, "coldatoms.jl_b200", "julia", "NeutralAtomsB200.jl")).read()
    src += open(os.path.join(ROOT, "tests", "julia_parity.jl")).read()   # the off-box parity script binds ca_rydberg1_run itself
    protos = _c_prototypes()
    assert all(not k.startswith("?") for ks in protos.values() for k in ks), protos
    checked = 0
    for m in re.finditer(r"ccall\(\s*((?:\w+\.)?sym\(:(\w+)\)|fp)\s*,\s*(\w+)\s*,\s*\(", src):
        i, depth = m.end(), 1
        while depth:   # the argument-type tuple
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        kinds = [_julia_kind(t) for t in _split_top(src[m.end():i - 1]) if t.strip()]
        assert all(not k.startswith("?") for k in kinds), kinds
        if m.group(2):
            names = [m.group(2)]
        else:   # fp, h = entry(:ca_x, :ca_x_multi) earlier in the same function
            e = list(re.finditer(r"entry\(:(\w+),\s*:(\w+)\)", src[:m.start()]))[-1]
            names = [e.group(1), e.group(2)]
        for name in names:
            assert name in protos, name
            assert kinds == protos[name], (name, kinds, protos[name])
            checked += 1
        assert (m.group(3) == "Cstring") == (names == ["ca_last_error"])   # every other entry returns the int status
    assert checked >= 19


def test_julia_struct_constructor_calls_have_a_matching_arity():
    """Julia's default constructor takes exactly one argument per field: every `CaXxx(...)` call in the shim must have that arity, or the
    arity of an outer constructor the shim defines (`CaModel(a::Vararg{Any,N}) = ...`)."""
    src = open(os.path.join(ROOT, "coldatoms.jl_b200", "julia", "NeutralAtomsB200.jl")).read()
    nfields = {k: len(v) for k, v in _julia_structs().items()}
    outer = {k: set() for k in nfields}
    for m in re.finditer(r"^(Ca\w+)\(\w+::Vararg\{Any,\s*(\d+)\}\)\s*=", src, re.M):
        outer[m.group(1)].add(int(m.group(2)))
    assert outer["CaModel"], "the 29-argument constructor of the cz_model.jl parameterisation is missing"
    calls = 0
    for m in re.finditer(r"(?<![\w{.:])(Ca(?:Beam|Model|Ode|Stats))\(", src):
        line_start = src.rfind("\n", 0, m.start()) + 1
        if re.match(r"\s*(?:mutable )?struct ", src[line_start:m.start() + 1]) or "::Vararg" in src[m.end():m.end() + 20]:
            continue
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1, "[": 1, "]": -1}.get(src[i], 0)
            i += 1
        args = [a for a in _split_top(src[m.end():i - 1].replace("[", "(").replace("]", ")")) if a.strip()]
        if m.group(1) == "CaStats" and not args:
            continue   # inner zero-argument constructor
        n = len(args)
        if any(a.strip().endswith("...") for a in args):
            continue   # splatted call (the outer constructor itself)
        assert n == nfields[m.group(1)] or n in outer[m.group(1)], (m.group(1), n, src[m.start():m.start() + 80])
        calls += 1
    assert calls >= 9


def test_python_mirror_call_sites_match_the_header(pkg):
    """Every `load().ca_xxx(...)` call in _lib.py against the header: arity, and the kind of each argument as far as the expression shows
    it (explicit ctypes scalar wrappers; pointers come from byref / _p / data_as / cast / None / array variables)."""
    src = open(os.path.join(ROOT, "coldatoms.jl_b200", "_lib.py")).read()
    protos = _c_prototypes()
    scalar = {"C.c_int32(": "i32", "C.c_int(": "i32", "C.c_int64(": "i64", "C.c_longlong(": "i64", "C.c_uint64(": "u64", "C.c_ulonglong(": "u64",
              "C.c_double(": "f64", "int(": "i32"}
    seen = set()
    for m in re.finditer(r"\.(ca_[a-z0-9_]+)\(", src):
        name = m.group(1)
        if src[m.start() - 4:m.start()] == "_abi":
            continue   # a ctypes struct constructor (_abi.ca_stats()), not a library call
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1, "[": 1, "]": -1}.get(src[i], 0)
            i += 1
        args = [a.strip() for a in _split_top(src[m.end():i - 1].replace("[", "(").replace("]", ")")) if a.strip()]
        assert name in protos, name
        assert len(args) == len(protos[name]), (name, len(args), len(protos[name]))
        for a, kind in zip(args, protos[name]):
            got = next((k for pre, k in scalar.items() if a.startswith(pre)), "ptr")
            assert got == kind, (name, a, got, kind)
        seen.add(name)
    assert seen >= set(pkg._lib.EXPORTS) - {"ca_version", "ca_multi_context"}, set(pkg._lib.EXPORTS) - seen
