#!/usr/bin/env python
"""Compiler-free check of the Fortran shim (fortran/wm_shim_modules.f90) against the reference's module interfaces.

No Fortran compiler exists in this image, so the only check of the drop-in claim that can run here is textual:
for every public module procedure the shim replaces, parse `subroutine name(args)` plus the declarations of its dummy
arguments out of the reference source and out of the shim and compare

  * module name, procedure name,
  * number, order and names of the dummy arguments,
  * type (integer / real(8) / external procedure) and rank (scalar / array) of every dummy argument,
  * intent -- equal, or one of the documented widenings (the shim may declare `intent(inout)` / `intent(in)` where the
    reference says `intent(out)` / `intent(inout)` for arrays it keeps on the device; listed in ALLOWED_INTENT).

    python scripts/check_shim_interfaces.py [/root/reference]        -> exit code 0 = every interface matches

With --json the parsed reference interfaces are written to tests/golden/ref_interfaces.json, so that the same comparison
runs in the CPU test suite on machines where /root/reference does not exist.
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_FILES = {
    "particle": "common/particle.f90", "field": "common/field.f90", "sort": "common/sort.f90",
    "mom_calc": "common/mom_calc.f90", "boundary_periodic": "common/boundary_periodic.f90",
    "boundary_shock": "proj/shock/boundary_shock.f90", "boundary_reconnection": "proj/reconnection/boundary_reconnection.f90",
}
# (module, procedure, argument): shim intent -> reference intent pairs that are deliberate (fortran/wm_shim_modules.f90 header)
ALLOWED_INTENT = {
    ("sort", "sort__bucket", "np2"): ("inout", "in"),   # the shim refreshes the host copy of np2 at a sync
}


def _join_continuations(text):
    out, cur = [], ""
    for raw in text.splitlines():
        line = raw.split("!")[0].rstrip() if not raw.lstrip().startswith("!$") else ""
        if not line.strip():
            continue
        s = line.strip()
        if s.startswith("&"):
            s = s[1:].lstrip()
        if s.endswith("&"):
            cur += s[:-1].rstrip() + " "
            continue
        out.append(cur + s)
        cur = ""
    return out


def _split_top(s):
    """split at commas that are not inside parentheses"""
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch == "(":
            depth += 1
        elif ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


def parse_module_procedures(text):
    """{module: {procedure: [(argname, type, rank, intent)]}} for the module-level subroutines (not interface bodies)."""
    lines = _join_continuations(text)
    mods, mod, proc, depth_iface = {}, None, None, 0
    for ln in lines:
        low = ln.lower().strip()
        m = re.match(r"module\s+(\w+)\s*$", low)
        if m and not low.startswith("module procedure"):
            mod = m.group(1)
            mods[mod] = {}
            continue
        if re.match(r"(abstract\s+)?interface\b", low):
            depth_iface += 1
            continue
        if re.match(r"end\s*interface", low):
            depth_iface -= 1
            continue
        if depth_iface:
            continue
        m = re.match(r"subroutine\s+(\w+)\s*\((.*)\)\s*$", low)
        if m and mod is not None and proc is None:
            proc = dict(name=m.group(1), args=[a.strip() for a in m.group(2).split(",") if a.strip()], decl={})
            continue
        if proc is not None:
            if re.match(r"end\s*subroutine", low):
                mods[mod][proc["name"]] = [(a,) + proc["decl"].get(a, ("?", "?", "?")) for a in proc["args"]]
                proc = None
                continue
            m = re.match(r"(integer|real\s*\(\s*8\s*\)|real\s*\(\s*c_double\s*\)|integer\s*\(\s*c_int32_t\s*\)|logical|external)\s*(.*?)(::)?\s*(.*)$", low)
            if m and ("::" in low or m.group(1) == "external"):
                typ = re.sub(r"\s+", "", m.group(1))
                typ = {"real(c_double)": "real(8)", "integer(c_int32_t)": "integer"}.get(typ, typ)
                attrs, names = low.split("::", 1) if "::" in low else (low, low[len("external"):])
                im = re.search(r"intent\s*\(\s*(\w+)\s*\)", attrs)
                intent = im.group(1) if im else ("external" if typ == "external" else "none")
                dim_attr = "dimension" in attrs
                for item in _split_top(names):
                    nm = re.match(r"(\w+)\s*(\(.*\))?", item.strip())
                    if not nm:
                        continue
                    rank = "array" if (nm.group(2) or dim_attr) else "scalar"
                    if nm.group(1) in proc["args"]:
                        proc["decl"][nm.group(1)] = (typ, rank, intent)
    return mods


def reference_interfaces(ref_root):
    out = {}
    for mod, rel in REF_FILES.items():
        with open(os.path.join(ref_root, rel)) as f:
            parsed = parse_module_procedures(f.read())
        pub = {k: v for k, v in parsed[mod].items() if k.startswith(mod + "__")}   # private helpers (ele_cur, cgm) stay inside
        out[mod] = dict(file=rel, procedures={k: [list(a) for a in v] for k, v in pub.items()})
    return out


def shim_interfaces():
    with open(os.path.join(ROOT, "fortran", "wm_shim_modules.f90")) as f:
        return parse_module_procedures(f.read())


def compare(ref, shim):
    problems = []
    for mod, info in ref.items():
        if mod not in shim:
            problems.append("module %s missing in the shim" % mod)
            continue
        for proc, rargs in info["procedures"].items():
            if proc not in shim[mod]:
                problems.append("%s: procedure %s missing in the shim" % (mod, proc))
                continue
            sargs = shim[mod][proc]
            if [a[0] for a in rargs] != [a[0] for a in sargs]:
                problems.append("%s: argument list differs: reference %s, shim %s" % (proc, [a[0] for a in rargs], [a[0] for a in sargs]))
                continue
            for (n, rt, rr, ri), (_, st, sr, si) in zip(rargs, sargs):
                if rt == "?" and st == "external":
                    continue   # procedure arguments: an interface block in the reference, `external` in the shim
                if (rt, rr) != (st, sr):
                    problems.append("%s(%s): reference %s %s, shim %s %s" % (proc, n, rt, rr, st, sr))
                if ri != si and ALLOWED_INTENT.get((mod, proc, n)) != (si, ri):
                    problems.append("%s(%s): intent reference %s, shim %s" % (proc, n, ri, si))
        extra = set(shim[mod]) - set(info["procedures"])
        if extra:
            problems.append("%s: the shim exports procedures the reference does not have: %s" % (mod, sorted(extra)))
    return problems


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    ref_root = args[0] if args else "/root/reference"
    gold = os.path.join(ROOT, "tests", "golden", "ref_interfaces.json")
    if os.path.isdir(ref_root):
        ref = reference_interfaces(ref_root)
        if "--json" in sys.argv:
            with open(gold, "w") as f:
                json.dump(ref, f, indent=1)
            print("wrote", gold)
    else:
        ref = json.load(open(gold))
    shim = shim_interfaces()
    problems = compare(ref, shim)
    n = sum(len(v["procedures"]) for v in ref.values())
    for p in problems:
        print("MISMATCH:", p)
    print("%d procedures of %d modules checked, %d mismatches" % (n, len(ref), len(problems)))
    return 1 if problems else 0


if __name__ == "__main__":
    sys.exit(main())
