#!/usr/bin/env python
"""The Fortran interface block of fortran/wm_cabi.f90 against the C prototypes of include/wumingpic2d.h.

No Fortran compiler in this image, and a compiler would not catch this class of error anyway: an `interface`
block is a promise about a C function that nobody checks -- a missing VALUE, an int32 where the C side takes int64, a swapped
argument compile and link and then corrupt memory at run time.  This script parses the interface block with numpy.f2py's
Fortran parser (crackfortran) and the header with a small prototype parser and compares, argument by argument:

    C  `wm_ctx *` / `const wm_ctx *` / `void *`          <->  type(c_ptr), value   (`void *` also: a buffer by reference)
    C  `wm_ctx **`                                        <->  type(c_ptr), intent(out)          (by reference)
    C  `const wm_config *`                                <->  type(wm_config), intent(in)       (by reference)
    C  `double *` / `const double *`                      <->  real(c_double), dimension(*)      (const <-> intent(in))
    C  `int32_t *` / `const int32_t *`                    <->  integer(c_int32_t), dimension(*)
    C  `char *` / `const char *`                          <->  character(kind=c_char), dimension(n)
    C  `double` / `int32_t` / `int64_t` / `size_t`        <->  real(c_double) / integer(c_int32_t / c_int64_t / c_size_t), value
    C  return `int`                                       <->  integer(c_int) result

and every call of a bound function in the shim sources passes as many actual arguments as its interface has dummies; and
`type, bind(C) :: wm_config` mirrors `struct wm_config` member by member (names, order, kinds, array lengths).

    python scripts/check_cabi_binding.py        exit code 0 = every bound function matches its prototype
"""
import contextlib
import io
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "wumingpic2d.h")
CABI = os.path.join(ROOT, "fortran", "wm_cabi.f90")


def c_prototypes(text=None):
    """{name: (return type, [(type, name), ...])} of every `wm_*` function declared in the header"""
    text = text if text is not None else open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    out = {}
    for m in re.finditer(r"\b((?:const\s+)?[A-Za-z_][A-Za-z0-9_]*(?:\s*\*)*)\s*\b(wm_[A-Za-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1), m.group(2), m.group(3)
        lst = []
        args = " ".join(args.split())
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                am = re.match(r"^(.*?)([A-Za-z_][A-Za-z0-9_]*)(\[[^\]]*\])?$", a)
                typ, an, arr = am.group(1).strip(), am.group(2), am.group(3)
                typ = re.sub(r"\s*\*\s*", "*", typ)
                typ = re.sub(r"\s+", " ", typ)
                if arr:
                    typ += "*"
                lst.append((typ, an))
        out[name] = (" ".join(ret.split()).replace(" *", "*"), lst)
    return out


def fortran_bindings(path=CABI, text=None):
    """{C name: (result kind, [(arg name, class), ...])} of the bind(C) functions of the interface block"""
    import numpy.f2py.crackfortran as cf
    cf.verbose = 0
    src = text if text is not None else open(path).read()
    tmp = None
    if text is not None:
        import tempfile
        tmp = tempfile.NamedTemporaryFile("w", suffix=".f90", delete=False)
        tmp.write(text)
        tmp.close()
        path = tmp.name
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(buf):
        blocks = cf.crackfortran([path])
    if tmp:
        os.unlink(tmp.name)
    # crackfortran drops the bind(C, name=...) clause: take the C names from the source text, keyed by the Fortran name
    cname = {m.group(1).lower(): m.group(2) for m in
             re.finditer(r"function\s+(\w+)\s*\([^)]*\)\s*bind\s*\(\s*C\s*,\s*name\s*=\s*'([^']+)'\s*\)", src, flags=re.I)}
    out = {}
    for mod in blocks:
        for b in mod.get("body", []):
            if b.get("block") != "interface":
                continue
            for f in b["body"]:
                if f.get("block") != "function" or f["name"].lower() not in cname:
                    continue
                args = []
                for a in f["args"]:
                    v = f["vars"][a]
                    kind = (v.get("kindselector") or {}).get("kind")
                    ts = v.get("typespec")
                    if ts == "type":
                        kind = v.get("typename")
                    if ts == "character":
                        kind = (v.get("charselector") or {}).get("kind", "c_char")
                    args.append((a, dict(typespec=ts, kind=kind, value="value" in (v.get("attrspec") or []),
                                         array=bool(v.get("dimension")), intent=(v.get("intent") or [None])[0])))
                r = f["vars"][f.get("result", f["name"])]
                out[cname[f["name"].lower()]] = ((r.get("typespec"), (r.get("kindselector") or {}).get("kind")), args)
    return out


SCALARS = {"double": ("real", "c_double"), "int32_t": ("integer", "c_int32_t"), "int64_t": ("integer", "c_int64_t"),
           "size_t": ("integer", "c_size_t"), "int": ("integer", "c_int")}


def match(ctype, f):
    """does the Fortran dummy `f` pass what the C parameter `ctype` expects?  Returns a problem string or None."""
    const = ctype.startswith("const ")
    base = ctype[6:] if const else ctype
    if base == "void*" and not f["value"] and (f["array"] or f["typespec"] == "character"):
        return None  # a buffer passed by reference (the 128-byte NCCL id)
    if base in ("wm_ctx*", "void*", "wm_loopback*"):
        ok = f["typespec"] == "type" and f["kind"] == "c_ptr" and f["value"] and not f["array"]
        return None if ok else "expects an opaque pointer by value: type(c_ptr), value"
    if base == "wm_ctx**":
        ok = f["typespec"] == "type" and f["kind"] == "c_ptr" and not f["value"]
        return None if ok else "expects the address of a pointer: type(c_ptr) without value"
    if base == "wm_config*":
        ok = f["typespec"] == "type" and f["kind"] == "wm_config" and not f["value"]
        return None if ok else "expects the address of a wm_config: type(wm_config) without value"
    if base == "char*":
        ok = f["typespec"] == "character" and not f["value"]
        return None if ok else "expects a character buffer by reference"
    if base.endswith("*") and base[:-1] in SCALARS:
        ts, kd = SCALARS[base[:-1]]
        if not (f["typespec"] == ts and f["kind"] == kd and not f["value"]):
            return "expects %s by reference: %s(%s) without value" % (base, ts, kd)
        if const and f["intent"] not in ("in", None):
            return "the C side only reads it (const): intent(in) expected, found intent(%s)" % f["intent"]
        if not const and f["intent"] == "in":
            return "the C side writes it (no const): intent(in) is a lie"
        return None
    if base in SCALARS:
        ts, kd = SCALARS[base]
        ok = f["typespec"] == ts and f["kind"] == kd and f["value"] and not f["array"]
        return None if ok else "expects %s by value: %s(%s), value" % (base, ts, kd)
    return "no rule for C type %r" % ctype


def compare(protos, binds):
    problems = []
    for name, (res, fargs) in sorted(binds.items()):
        if not name.startswith("wm_"):
            continue
        if name not in protos:
            problems.append("%s: bound in wm_cabi.f90 but not declared in wumingpic2d.h" % name)
            continue
        ret, cargs = protos[name]
        if ret == "int" and res != ("integer", "c_int"):
            problems.append("%s: returns int, the interface says %s(%s)" % (name, res[0], res[1]))
        if ret in ("const char*",) and not (res[0] == "type"):
            problems.append("%s: returns a C string, the interface must return type(c_ptr)" % name)
        if len(cargs) != len(fargs):
            problems.append("%s: %d C parameters, %d Fortran dummies" % (name, len(cargs), len(fargs)))
            continue
        for (ct, cn), (fn, fa) in zip(cargs, fargs):
            p = match(ct, fa)
            if p:
                problems.append("%s: argument %s (C: %s %s): %s" % (name, fn, ct, cn, p))
    return problems


def call_sites(files=None):
    """[(file, line, C function, number of actual arguments)] of every call of a bound function in the shim sources"""
    files = files or [CABI, os.path.join(ROOT, "fortran", "wm_shim_modules.f90")]
    out = []
    for path in files:
        src = open(path).read()
        # join continuation lines, drop comments
        lines = []
        for ln in src.split("\n"):
            ln = re.sub(r"'[^']*'", "''", ln)       # string literals (they quote the function names in error messages)
            lines.append(ln.split("!")[0])          # comments
        txt = "\n".join(lines)
        txt = re.sub(r"&\s*\n\s*&?", " ", txt)
        for m in re.finditer(r"\b(wm_[a-z0-9_]+)\s*\(", txt, flags=re.I):
            name = m.group(1)
            head = txt[max(0, m.start() - 40):m.start()]
            if re.search(r"(function|subroutine)\s+$", head, flags=re.I) or "name=" in head[-12:]:
                continue  # a declaration, not a call
            depth, i, nargs, seen = 1, m.end(), 0, False
            while i < len(txt) and depth:
                ch = txt[i]
                if ch == "(":
                    depth += 1
                elif ch == ")":
                    depth -= 1
                elif ch == "," and depth == 1:
                    nargs += 1
                elif not ch.isspace():
                    seen = True
                i += 1
            out.append((os.path.basename(path), txt.count("\n", 0, m.start()) + 1, name, nargs + 1 if seen else 0))
    return out


def check_calls(binds, sites=None):
    problems = []
    for f, ln, name, n in sites if sites is not None else call_sites():
        if name in binds and n != len(binds[name][1]):
            problems.append("%s:%d: %s called with %d arguments, its interface has %d" % (f, ln, name, n, len(binds[name][1])))
    return problems


def c_struct_members(text=None, name="wm_config"):
    """[(C type, member, array length or None)] of `struct name` in the header, in declaration order"""
    text = text if text is not None else open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    m = re.search(r"typedef\s+struct\s+%s\s*\{(.*?)\}\s*%s\s*;" % (name, name), text, flags=re.S)
    out = []
    for st in m.group(1).split(";"):
        st = " ".join(st.split())
        if not st:
            continue
        typ, rest = st.split(" ", 1)
        for mem in rest.split(","):
            mm = re.match(r"\s*(\w+)\s*(?:\[\s*(\w+)\s*\])?", mem)
            out.append((typ, mm.group(1), mm.group(2)))
    return out


def fortran_type_components(text=None, name="wm_config"):
    """[(typespec, kind, component, array length or None)] of `type, bind(C) :: name` in wm_cabi.f90, in declaration order"""
    text = text if text is not None else open(CABI).read()
    m = re.search(r"^\s*type\s*,\s*bind\s*\(\s*C\s*\)\s*::\s*%s\s*$(.*?)^\s*end\s+type" % name, text, flags=re.S | re.M | re.I)
    out = []
    for ln in m.group(1).split("\n"):
        ln = ln.split("!")[0].strip()
        if not ln:
            continue
        dm = re.match(r"(integer|real)\s*\(\s*(\w+)\s*\)\s*::\s*(.*)$", ln, flags=re.I)
        for comp in re.findall(r"(\w+)\s*(?:\(\s*(\w+)\s*\))?\s*(?:,|$)", dm.group(3)):
            out.append((dm.group(1).lower(), dm.group(2).lower(), comp[0], comp[1] or None))
    return out


def compare_struct(cm=None, fm=None):
    """member by member: the same names in the same order, interoperable kinds, the same array lengths"""
    cm = cm if cm is not None else c_struct_members()
    fm = fm if fm is not None else fortran_type_components()
    problems = []
    if len(cm) != len(fm):
        problems.append("wm_config: %d C members, %d Fortran components" % (len(cm), len(fm)))
    for (ct, cn, cl), (ft, fk, fn, fl) in zip(cm, fm):
        if cn.lower() != fn.lower():
            problems.append("wm_config: member %s of the C struct is component %s in Fortran (order matters: bind(C))" % (cn, fn))
        elif SCALARS.get(ct) != (ft, fk):
            problems.append("wm_config%%%s: C %s, Fortran %s(%s)" % (fn, ct, ft, fk))
        elif (cl or "").lower() != (fl or "").lower():
            problems.append("wm_config%%%s: array length %s in C, %s in Fortran" % (fn, cl, fl))
    return problems


def main():
    protos, binds = c_prototypes(), fortran_bindings()
    probs = compare(protos, binds) + check_calls(binds) + compare_struct()
    print("%d functions bound in fortran/wm_cabi.f90, %d declared in include/wumingpic2d.h, %d problems"
          % (len([n for n in binds if n.startswith("wm_")]), len(protos), len(probs)))
    for p in probs:
        print("  " + p)
    return 1 if probs else 0


if __name__ == "__main__":
    sys.exit(main())
