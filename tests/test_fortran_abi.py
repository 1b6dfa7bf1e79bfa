"""The Fortran layers cannot be compiled here (no Fortran compiler in the image), so the
binding surface is checked from the C side: every `bind(C, name=...)` in
nka_b200/fortran/nka_b200_c.F90 must be exported by the library, declared in include/*.h with
the same number of arguments, and pass scalars by value exactly where the header does."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
F90 = os.path.join(ROOT, "nka_b200", "fortran", "nka_b200_c.F90")
F90_EXAMPLE = os.path.join(ROOT, "nka_b200", "fortran", "nka_example_c.F90")


def _fortran_interfaces():
    text = open(F90).read() + open(F90_EXAMPLE).read()
    text = text.replace("&\n", " ")
    out = {}
    pat = re.compile(r"(?:function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)(.*?)end (?:function|subroutine)",
                     re.S | re.I)
    for m in pat.finditer(text):
        args = [a.strip() for a in m.group(2).split(",") if a.strip()]
        body = m.group(4)
        by_value = set()
        for line in body.splitlines():
            if "value" in line.lower() and "::" in line:
                by_value.update(n.strip() for n in line.split("::")[1].split(","))
        out[m.group(3)] = {"args": args, "by_value": by_value}
    return out


def _c_declarations():
    decls = {}
    for fn in os.listdir(os.path.join(ROOT, "include")):
        text = open(os.path.join(ROOT, "include", fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"\b(nka_\w+)\s*\(([^;{]*?)\)\s*;", text, re.S):
            params = m.group(2).strip()
            params = re.sub(r"\([^()]*\)", "", params)        # drop nested (fn-pointer) parens
            n = 0 if params in ("", "void") else len([p for p in params.split(",") if p.strip()])
            decls[m.group(1)] = {"n": n, "params": [p.strip() for p in m.group(2).split(",")]}
    return decls


def test_every_fortran_binding_is_exported_and_declared():
    from nka_b200 import build
    path = build.build_library()
    exported = set(subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True,
                                  check=True).stdout.split())
    ifaces = _fortran_interfaces()
    decls = _c_declarations()
    assert len(ifaces) >= 55
    for name, info in ifaces.items():
        assert name in exported, name
        assert name in decls, name
        if name == "nka_init":
            continue
        assert len(info["args"]) == decls[name]["n"], (name, info["args"], decls[name])


def test_scalars_by_value_arrays_by_address():
    """A C `double`/`int`/`size_t`/handle parameter must be `value` in Fortran; a `double *` array
    or the 128-byte id must not be."""
    ifaces = _fortran_interfaces()
    decls = _c_typed_declarations()                 # parameters split on top-level commas only
    for name, info in ifaces.items():
        for farg, cparam in zip(info["args"], decls[name]["params_raw"]):
            is_array = ("double *" in cparam and "NKA" not in cparam and "(*" not in cparam) or "id128" in cparam
            if name in ("nka_accel_update_dev",) and "f_dev" in farg:
                is_array = False          # device address travels by value in a type(c_ptr)
            assert (farg in info["by_value"]) == (not is_array), (name, farg, cparam)


def test_fortran_modules_keep_reference_names():
    """Module, type and procedure names of the three reference flavours are kept."""
    f95 = open(os.path.join(ROOT, "nka_b200", "fortran", "F95", "nka_type.F90")).read()
    for name in ("nka_init", "nka_delete", "nka_set_vec_tol", "nka_defined", "nka_vec_len", "nka_num_vec",
                 "nka_max_vec", "nka_vec_tol", "nka_real_kind", "nka_accel_update", "nka_relax", "nka_restart"):
        assert re.search(r"public ::.*\b%s\b" % name, f95), name       # src-F95/nka_type.F90:205-207
    assert "module nka_type" in f95 and "type, public :: nka" in f95
    f08 = open(os.path.join(ROOT, "nka_b200", "fortran", "F08", "nka_type.F90")).read()
    for name in ("init", "set_vec_tol", "set_dot_prod", "vec_len", "num_vec", "max_vec", "vec_tol",
                 "accel_update", "relax", "restart", "defined"):                 # src-F08/nka_type.F90:169-181
        assert re.search(r"(procedure|generic)\s*(,\s*private)?\s*::\s*%s\b" % name, f08), name
    vec = open(os.path.join(ROOT, "nka_b200", "fortran", "F08-vector", "gpu_vector_type.F90")).read()
    for name in ("clone1", "clone2", "copy_", "setval", "scale", "update1_", "update2_", "update3_", "update4_",
                 "dot_", "norm2"):                                               # vector_class.F90:93-108
        assert re.search(r"procedure :: %s\b" % name, vec), name
    assert "type, extends(vector), public :: gpu_vector" in vec


# ---------------------------------------------------------------------------------------------
# Type-exact check.  For every bind(C) interface the Fortran dummy arguments and result are
# translated to the C types ISO_C_BINDING says they interoperate with (Fortran 2003 15.2.2-15.3.6:
# integer(c_int) <-> int, integer(c_size_t) <-> size_t, integer(c_long_long) <-> long long,
# real(c_double) <-> double, type(c_ptr) <-> void * (any object pointer), character(kind=c_char)
# <-> char; VALUE = by value, otherwise by address), then
#   (1) compared with the header's parameter types one by one, and
#   (2) written out as C function-pointer typedefs, initialised from the header's own
#       declarations, and compiled with gcc -Werror: a c_int bound to a size_t parameter, a missing
#       VALUE, a wrong result kind or a wrong argument count does not compile.
# ---------------------------------------------------------------------------------------------
_F_SCALAR = {"integer(c_int)": "int", "integer(c_size_t)": "size_t", "integer(c_long_long)": "long long",
             "real(c_double)": "double", "type(c_ptr)": "ptr", "type(c_funptr)": "fnptr",
             "character(kind=c_char)": "char"}
_HANDLES = ("NKA", "NKAVEC", "NKASYS")


def _fortran_typed_interfaces():
    text = open(F90).read() + open(F90_EXAMPLE).read()
    text = re.sub(r"&\s*\n\s*", " ", text)
    out = {}
    pat = re.compile(r"(function|subroutine)\s+(\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)\s*(?:result\((\w+)\))?(.*?)end (?:function|subroutine)",
                     re.S | re.I)
    for m in pat.finditer(text):
        kind, args, cname, resname, body = m.group(1).lower(), m.group(3), m.group(4), m.group(5), m.group(6)
        args = [a.strip() for a in args.split(",") if a.strip()]
        decl = {}
        for line in body.splitlines():
            line = line.split("!")[0]
            if "::" not in line or line.strip().lower().startswith("import"):
                continue
            spec, names = line.split("::")
            spec = spec.strip()
            base = spec.split(",")[0].strip().replace(" ", "").lower()
            base = {"character(kind=c_char)": "character(kind=c_char)"}.get(base, base)
            attrs = [a.strip().lower() for a in spec.split(",")[1:]]
            for nm in re.findall(r"(\w+)\s*(\([^)]*\))?", names):
                if nm[0]:
                    decl[nm[0]] = {"base": base, "value": "value" in attrs,
                                   "array": bool(nm[1]) or any(a.startswith("dimension") for a in attrs)}
        params = []
        for a in args:
            d = decl[a]
            c = _F_SCALAR[d["base"]]
            if d["value"]:
                assert not d["array"], (cname, a)
                params.append(c)                       # by value
            else:
                params.append(c + "*")                 # by address (arrays and non-VALUE scalars alike)
        res = "void"
        if kind == "function":
            res = _F_SCALAR[decl[resname or m.group(2)]["base"]]
        out[cname] = {"params": params, "result": res}
    return out


def _canon_c_type(p):
    """'const double *host' -> 'double*', 'NKA' -> 'ptr', 'double ms[5]' -> 'double*', 'int on' -> 'int'."""
    p = p.strip()
    if "(*" in p:
        return "fnptr"
    is_array = "[" in p
    p = re.sub(r"\[[^\]]*\]", "", p)
    stars = p.count("*")
    p = p.replace("*", " ").replace("const", " ")
    words = p.split()
    known = {"int", "double", "size_t", "void", "char", "unsigned", "long", "nka_state_view"} | set(_HANDLES)
    if len(words) > 1 and words[-1] not in known:
        words = words[:-1]                             # drop the parameter name
    base = " ".join(words)
    if base in _HANDLES:
        base, stars = "void", stars + 1                # typedef struct ... *NKA
    if is_array:
        stars += 1
    return base + "*" * stars


def _c_typed_declarations():
    decls = {}
    for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
        text = open(os.path.join(ROOT, "include", fn)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        for m in re.finditer(r"(?:^|\n)\s*(?:extern\s+)?([\w ]+?[\s\*]+)(nka_\w+)\s*\((.*?)\)\s*;", text, re.S):
            ret, name, params = m.group(1), m.group(2), m.group(3)
            depth, cur, parts = 0, "", []
            for ch in params:                          # split on top-level commas only
                if ch == "(":
                    depth += 1
                if ch == ")":
                    depth -= 1
                if ch == "," and depth == 0:
                    parts.append(cur)
                    cur = ""
                else:
                    cur += ch
            parts.append(cur)
            parts = [] if params.strip() in ("", "void") else parts
            decls[name] = {"params_raw": [p.strip() for p in parts], "params": [_canon_c_type(p) for p in parts],
                           "result": _canon_c_type(ret + " x").replace(" x", "") if ret.strip() != "void" else "void",
                           "result_raw": ret.strip()}
    return decls


def _interoperates(f, c):
    """Fortran-side C type f (from ISO_C_BINDING) against the header's canonical type c."""
    if f == c:
        return True
    if f == "ptr":                                     # type(c_ptr), value: any object pointer
        return c.endswith("*") and c != "fnptr"
    if f == "fnptr" or c == "fnptr":
        return f in ("fnptr", "ptr") and c == "fnptr"
    if f == "long long" and c == "unsigned long long":
        return True                                    # Fortran has no unsigned kinds: same size and representation
    if f == "long long*" and c == "unsigned long long*":
        return True
    if f == "char*" and c in ("void*", "char*"):
        return True                                    # the 128-byte NCCL id / version string: raw bytes
    if f == "ptr*":
        return False
    return False


def test_fortran_kinds_match_c_parameter_types_exactly():
    fi = _fortran_typed_interfaces()
    cd = _c_typed_declarations()
    assert len(fi) >= 55
    checked = 0
    for name, f in fi.items():
        c = cd[name]
        assert len(f["params"]) == len(c["params"]), (name, f["params"], c["params_raw"])
        for k, (fp, cp) in enumerate(zip(f["params"], c["params"])):
            assert _interoperates(fp, cp), (name, k, fp, c["params_raw"][k])
            checked += 1
        assert _interoperates(f["result"], c["result"]) or (f["result"] == c["result"] == "void"), \
            (name, f["result"], c["result_raw"])
    assert checked >= 120


def test_fortran_binding_shim_compiles_against_the_headers(tmp_path):
    """The Fortran view of every entry point, as C function-pointer types, initialised from the
    headers' declarations: gcc -Werror rejects any mismatch in arity, scalar width, VALUE-ness or
    result kind (pointer parameters take the header's own pointer type when the Fortran side
    passes an address, so what is checked there is pointer-versus-scalar)."""
    fi = _fortran_typed_interfaces()
    cd = _c_typed_declarations()
    scalar = {"int": "int", "size_t": "size_t", "long long": "long long", "double": "double"}
    lines = ['#include <stddef.h>', '#include "nonlinear_krylov_accelerator.h"', '#include "nka_b200.h"',
             '#include "nka_example.h"', ""]
    for name, f in sorted(fi.items()):
        c = cd[name]

        def ctype(ftype, raw):
            if ftype in scalar:
                # unsigned long long in the header: Fortran's c_long_long has the same width
                return "unsigned long long" if (ftype == "long long" and "unsigned" in raw) else scalar[ftype]
            raw_is_pointer = ("*" in raw) or ("[" in raw) or raw.split()[0] in _HANDLES
            if not raw_is_pointer:
                return "void *"                        # Fortran passes an address, the header wants a scalar: must not compile
            t = re.sub(r"\[[^\]]*\]", "", raw)
            if "(*" in t:                             # function pointer: the header's own type, name removed
                return re.sub(r"\(\*\s*\w+\)", "(*)", t)
            words = t.replace("*", " * ").split()
            if words[-1] != "*" and words[-1] not in _HANDLES and len(words) > 1:
                words = words[:-1]
            return " ".join(words) + (" *" if "[" in raw else "")
        params = ", ".join(ctype(fp, raw) for fp, raw in zip(f["params"], c["params_raw"])) or "void"
        res = "void" if f["result"] == "void" else ctype(f["result"], c["result_raw"] + " *" if False else c["result_raw"])
        if f["result"] == "ptr":
            res = c["result_raw"]
        lines.append("typedef %s (*ft_%s)(%s);" % (res, name, params))
        lines.append("static ft_%s chk_%s = %s;" % (name, name, name))
    lines.append("int nka_fortran_abi_shim_entries(void) { return %d; }" % len(fi))
    src = tmp_path / "fortran_abi_shim.c"
    src.write_text("\n".join(lines) + "\n")
    r = subprocess.run(["gcc", "-std=c11", "-Wall", "-Wno-unused-variable", "-Werror", "-Werror=incompatible-pointer-types",
                        "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "shim.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + "\n" + src.read_text()


def test_shim_check_really_rejects_a_width_mismatch(tmp_path):
    """The compile check has teeth: binding nka_init_ex's size_t length as an int must fail."""
    src = tmp_path / "bad.c"
    src.write_text('#include <stddef.h>\n#include "nka_b200.h"\n'
                   "typedef NKA (*ft)(int, int, double, int, void *);\nstatic ft chk = nka_init_ex;\n")
    r = subprocess.run(["gcc", "-std=c11", "-Werror=incompatible-pointer-types", "-Wno-unused-variable",
                        "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "bad.o")],
                       capture_output=True, text=True)
    assert r.returncode != 0


def test_fortran_sources_parse_with_f2py_crackfortran():
    """No Fortran compiler in the image; numpy's f2py carries a Fortran 90 parser (crackfortran).  It is not a
    compiler -- it does not check types or generic resolution -- but it does find unbalanced blocks, broken
    continuation lines and malformed declarations: every module must parse, and in the two interface modules it
    must see exactly the bind(C) procedures the ABI checks above work from."""
    import contextlib
    import glob
    import io
    from numpy.f2py import crackfortran
    fdir = os.path.join(ROOT, "nka_b200", "fortran")
    files = sorted(glob.glob(os.path.join(fdir, "*.F90")) + glob.glob(os.path.join(fdir, "*", "*.F90")))
    assert len(files) == 6
    want_modules = {"nka_b200_c.F90": "nka_b200_c", "nka_example_c.F90": "nka_example_c", "gpu_vector_type.F90": "gpu_vector_type"}
    for f in files:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(buf):
            blocks = crackfortran.crackfortran([f])
        mods = [b for b in blocks if b.get("block") == "module"]
        assert len(mods) == 1, (f, [b.get("block") for b in blocks])
        assert mods[0]["name"] == want_modules.get(os.path.basename(f), "nka_type"), f
        if os.path.basename(f) in ("nka_b200_c.F90", "nka_example_c.F90"):
            procs = []

            def walk(body):
                for b in body:
                    if b.get("block") in ("function", "subroutine"):
                        procs.append(b["name"])
                    walk(b.get("body", []))
            walk(mods[0]["body"])
            text = re.sub(r"&\s*\n\s*", " ", open(f).read())
            declared = re.findall(r"(?:function|subroutine)\s+(\w+)\s*\([^)]*\)\s*bind\(C", text, re.I)
            assert sorted(p.lower() for p in procs) == sorted(d.lower() for d in declared), f
