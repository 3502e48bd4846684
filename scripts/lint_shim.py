#!/usr/bin/env python
"""Identifier lint of the Fortran shim (fortran/*.f90), for want of a compiler: every name used in the executable
part of a procedure must be a dummy argument, a local declared in that procedure, a module variable / procedure / bound C
function of its module or of `wm_cabi`, an intrinsic, or an MPI name.  Catches the typo that would stop a maintainer's first
build; it is not a type checker.

    python scripts/lint_shim.py       exit code 0 = no unknown identifier
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = [os.path.join(ROOT, "fortran", "wm_cabi.f90"), os.path.join(ROOT, "fortran", "wm_shim_modules.f90")]

KNOWN = set("""
if then else elseif end endif do enddo call return stop write read print use only implicit none subroutine function module contains
integer real logical character type intent in out inout value target save parameter dimension allocatable pointer result bind
and or not true false eq ne lt le gt ge kind len trim adjustl int dble mod min max abs sqrt size present associated exit cycle while
select case default status import interface procedure external unit iostat transfer achar iachar huge tiny nint floor ceiling
get_environment_variable iso_c_binding intrinsic c_associated c_loc c_f_pointer c_null_ptr c_null_char c_char c_int c_int32_t
c_int64_t c_double c_ptr c_size_t c_funptr
mpi mpi_bcast mpi_comm_rank mpi_comm_size mpi_character mpi_integer mpi_comm_world
""".split())
DECL = re.compile(r"^\s*(integer|real|logical|character|double\s+precision|complex|type\s*\()", re.I)


def strip(src):
    out = []
    for ln in src.split("\n"):
        ln = re.sub(r"'[^']*'", "''", ln)
        out.append(ln.split("!")[0])
    txt = "\n".join(out)
    return re.sub(r"&\s*\n\s*&?", " ", txt)


def declared_names(stmt):
    """names introduced by one declaration statement"""
    rhs = stmt.split("::", 1)[1] if "::" in stmt else re.sub(DECL, "", stmt, count=1)
    names, depth, cur = [], 0, ""
    for ch in rhs + ",":
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            m = re.match(r"\s*([A-Za-z_]\w*)", cur)
            if m:
                names.append(m.group(1).lower())
            cur = ""
        else:
            cur += ch
    return names


def module_symbols(txt):
    """{module: (module-level names incl. procedures, types, bound functions, enumerators)}"""
    mods = {}
    for m in re.finditer(r"^\s*module\s+(\w+)\s*$(.*?)^\s*end\s+module\s+\1", txt, flags=re.S | re.M | re.I):
        name, body = m.group(1).lower(), m.group(2)
        head = re.split(r"^\s*contains\s*$", body, maxsplit=1, flags=re.M | re.I)[0]
        syms = set()
        types = {x.lower() for x in re.findall(r"^\s*type\b(?!\s*\()[^\n]*?(\w+)\s*$", head, flags=re.M | re.I)}
        head = re.sub(r"^\s*type\b(?!\s*\()[^\n]*\n.*?^\s*end\s+type\b[^\n]*$", "", head, flags=re.S | re.M | re.I)   # components are not module names
        head = re.sub(r"^\s*interface\b.*?^\s*end\s+interface\b[^\n]*$", "", head, flags=re.S | re.M | re.I)  # dummies of the bound functions neither
        syms |= types
        for st in head.split("\n"):
            if DECL.match(st) and "function" not in st.lower():
                syms |= set(declared_names(st))
        syms |= {x.lower() for x in re.findall(r"^\s*type\s*,?[^:\n]*::\s*(\w+)", head, flags=re.M | re.I)}
        syms |= {x.lower() for x in re.findall(r"^\s*(?:enumerator\s*::|integer\s*\([^)]*\)\s*,\s*parameter\s*::)\s*(\w+)", head, flags=re.M | re.I)}
        syms |= {x.lower() for x in re.findall(r"\b(?:function|subroutine)\s+(\w+)", body, flags=re.I)}
        mods[name] = (syms, body)
    return mods


def lint(files=FILES, texts=None):
    texts = texts or {f: open(f).read() for f in files}
    mods = {}
    for f, src in texts.items():
        for k, v in module_symbols(strip(src)).items():
            mods[k] = v + (os.path.basename(f),)
    problems = []
    for mname, (syms, body, fname) in mods.items():
        used = {u.lower() for u in re.findall(r"^\s*use\s+(\w+)", body, flags=re.M | re.I)}
        visible = set(syms)
        for u in used:
            if u in mods:
                visible |= mods[u][0]
        parts = re.split(r"^\s*contains\s*$", body, maxsplit=1, flags=re.M | re.I)
        if len(parts) < 2:
            continue
        for pm in re.finditer(r"^\s*(?:subroutine|function)\s+(\w+)\s*(\([^)]*\))?[^\n]*\n(.*?)^\s*end\s+(?:subroutine|function)\b", parts[1], flags=re.S | re.M | re.I):
            pname, args, pbody = pm.group(1).lower(), pm.group(2) or "", pm.group(3)
            local = {a.strip().lower() for a in args.strip("()").split(",") if a.strip()} | {pname}
            exe = []
            for st in pbody.split("\n"):
                if DECL.match(st):
                    local |= set(declared_names(st))
                    # dimension expressions and kinds of a declaration use names too (cfg%nys ...): check them as well
                    exe.append(st.split("::")[0] if "::" in st else "")
                    if "::" in st:
                        exe.append(" ".join(re.findall(r"\(([^()]*)\)", st.split("::", 1)[1])))
                elif re.match(r"^\s*use\b", st, flags=re.I):
                    continue
                else:
                    exe.append(st)
            code = "\n".join(exe)
            toks = {t.lower() for t in re.findall(r"(?<![%\w.])[A-Za-z_]\w*", code)}
            toks -= {t.lower() for t in re.findall(r"\.(\w+)\.", code)}
            unknown = sorted(t for t in toks if t not in local and t not in visible and t not in KNOWN
                             and not re.match(r"^[de]\d+$", t) and not re.match(r"^\d", t) and not re.fullmatch(r"\d+_\w+", t))
            unknown = [t for t in unknown if not re.search(r"\d_%s\b" % re.escape(t), code)]  # kind suffixes: 8_c_size_t
            if unknown:
                problems.append("%s: %s::%s uses undeclared %s" % (fname, mname, pname, unknown))
    return problems


def main():
    probs = lint()
    print("%d problems" % len(probs))
    for p in probs:
        print("  " + p)
    return 1 if probs else 0


if __name__ == "__main__":
    sys.exit(main())
