"""Row f4 of SURVEY.md section 8 without a Fortran compiler: the shim's module procedures must have the reference's names,
argument lists, types, ranks and intents (scripts/check_shim_interfaces.py; the reference's interfaces are frozen in
tests/golden/ref_interfaces.json by the same script run with --json against /root/reference)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_shim_matches_reference_interfaces():
    import check_shim_interfaces as C
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_interfaces.json")))
    assert sum(len(v["procedures"]) for v in ref.values()) == 31
    assert C.compare(ref, C.shim_interfaces()) == []


def test_checker_sees_a_swapped_argument():
    import check_shim_interfaces as C
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_interfaces.json")))
    txt = open(os.path.join(ROOT, "fortran", "wm_shim_modules.f90")).read()
    bad = txt.replace("subroutine sort__bucket(gp,up,cumcnt,np2,nxs,nxe)", "subroutine sort__bucket(up,gp,cumcnt,np2,nxs,nxe)", 1)
    assert bad != txt
    assert any("sort__bucket" in p for p in C.compare(ref, C.parse_module_procedures(bad)))
    bad = txt.replace("real(8), intent(in)    :: u0", "integer, intent(in)    :: u0", 1)
    assert bad != txt
    assert any("u0" in p for p in C.compare(ref, C.parse_module_procedures(bad)))


def test_golden_interfaces_are_current():
    """where the reference tree is present (the build container), the frozen interfaces must equal a fresh parse"""
    import check_shim_interfaces as C
    import pytest
    if not os.path.isdir("/root/reference"):
        pytest.skip("no reference tree on this machine")
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_interfaces.json")))
    assert C.reference_interfaces("/root/reference") == ref
