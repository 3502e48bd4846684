"""fortran/wm_cabi.f90's interface block against the prototypes of include/wumingpic2d.h (scripts/check_cabi_binding.py): an
interface block is an unchecked promise about a C function -- a missing VALUE or a wrong integer kind links fine and corrupts
memory at run time, and no Fortran compiler (absent here anyway) would notice.  Parsed with numpy.f2py's
crackfortran, so the shim has at least met A Fortran parser.  CPU only."""
import os
import sys

import pytest

pytestmark = pytest.mark.filterwarnings("ignore")   # (the parser warns about cfg%nye in dimension expressions)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_every_bound_function_matches_its_prototype():
    import check_cabi_binding as B
    protos, binds = B.c_prototypes(), B.fortran_bindings()
    bound = [n for n in binds if n.startswith("wm_")]
    assert len(bound) >= 26 and len(protos) >= 58
    # the procedures the shim modules call must all be bound
    for need in ("wm_create", "wm_particle__solv", "wm_field__fdtd_i", "wm_boundary__particle_x", "wm_boundary__particle_y",
                 "wm_sort__bucket", "wm_mom_calc__accl", "wm_mom_calc__nvt", "wm_boundary__mom", "wm_upload_particles_sorted",
                 "wm_download_particles", "wm_set_xrange", "wm_boundary__injection", "wm_host_register"):
        assert need in binds, need
    assert B.compare(protos, binds) == []


def test_the_checker_sees_the_classic_mistakes():
    import check_cabi_binding as B
    protos = B.c_prototypes()
    src = open(B.CABI).read()
    # (1) a scalar passed by reference instead of by value
    bad = src.replace("integer(c_int32_t), value :: nsteps", "integer(c_int32_t) :: nsteps", 1)
    assert bad != src
    assert any(p.startswith("wm_step:") for p in B.compare(protos, B.fortran_bindings(text=bad)))
    # (2) the wrong integer kind
    bad = src.replace("integer(c_int64_t), value :: n", "integer(c_int32_t), value :: n", 1)
    assert bad != src
    assert any(p.startswith("wm_append_particles:") for p in B.compare(protos, B.fortran_bindings(text=bad)))
    # (3) an output array declared intent(in)
    bad = src.replace("real(c_double), intent(inout)  :: gp(*)", "real(c_double), intent(in)  :: gp(*)", 1)
    if bad != src:
        assert any(p.startswith("wm_download_gp:") for p in B.compare(protos, B.fortran_bindings(text=bad)))


def test_shim_files_parse():
    """module and procedure structure of both shim files as a Fortran parser sees it"""
    import contextlib
    import io
    import numpy.f2py.crackfortran as cf
    cf.verbose = 0
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(buf):
        blocks = cf.crackfortran([os.path.join(ROOT, "fortran", "wm_shim_modules.f90")])
    mods = {b["name"]: [x["name"] for x in b["body"] if x.get("block") in ("subroutine", "function")] for b in blocks if b.get("block") == "module"}
    assert set(mods) == {"particle", "field", "sort", "boundary_periodic", "mom_calc", "boundary_reconnection", "boundary_shock"}
    assert mods["particle"] == ["particle__init", "particle__solv"]
    assert mods["sort"] == ["sort__init", "sort__bucket"]
    assert len(mods["boundary_periodic"]) == 7 and len(mods["boundary_shock"]) == 8
    assert sum(len(v) for v in mods.values()) == 31


def test_call_sites_pass_the_right_number_of_arguments():
    import check_cabi_binding as B
    binds = B.fortran_bindings()
    sites = [s for s in B.call_sites() if s[2] in binds]
    assert len(sites) >= 30
    assert B.check_calls(binds, sites) == []
    assert B.check_calls(binds, [("x.f90", 1, "wm_step", 1)]) != []


def test_shim_uses_only_declared_names():
    """scripts/lint_shim.py: every identifier in the executable part of the shim's 38 procedures is a dummy, a declared local, a
    module variable / procedure / bound function, an intrinsic or an MPI name"""
    import lint_shim as L
    assert L.lint() == []


def test_lint_sees_typos():
    import lint_shim as L
    texts = {f: open(f).read() for f in L.FILES}
    cabi, mods = L.FILES
    bad = dict(texts)
    bad[cabi] = texts[cabi].replace("    nstep_since_sync = 0\n  end subroutine wm_shim__download", "    nstep_since_sink = 0\n  end subroutine wm_shim__download")
    assert bad[cabi] != texts[cabi] and any("nstep_since_sink" in p for p in L.lint(texts=bad))
    bad = dict(texts)
    bad[mods] = texts[mods].replace("call wm_check(wm_sort__bucket(ctx), 'sort__bucket')", "call wm_check(wm_sort__buckets(ctx), 'sort__bucket')")
    assert bad[mods] != texts[mods] and any("wm_sort__buckets" in p for p in L.lint(texts=bad))
    bad = dict(texts)   # a local that lost its declaration (and must not be mistaken for the component cfg%nrank)
    bad[mods] = texts[mods].replace("    integer :: nerr, nrank, nsize\n    call MPI_COMM_RANK(ncomw_in, nrank, nerr)",
                                    "    integer :: nerr, nsize\n    call MPI_COMM_RANK(ncomw_in, nrank, nerr)", 1)
    assert bad[mods] != texts[mods] and any("nrank" in p for p in L.lint(texts=bad))


def test_wm_config_is_mirrored_member_by_member():
    """struct wm_config (include/wumingpic2d.h) == type, bind(C) :: wm_config (fortran/wm_cabi.f90) == api.WmConfig (ctypes):
    names, order, kinds, array lengths -- a mismatch would shift every member behind it"""
    import ctypes as C
    import re
    import check_cabi_binding as B
    cm, fm = B.c_struct_members(), B.fortran_type_components()
    assert len(cm) == 21 and B.compare_struct(cm, fm) == []
    swapped = list(fm)
    swapped[7], swapped[8] = swapped[8], swapped[7]
    assert B.compare_struct(cm, swapped) != []
    assert B.compare_struct(cm, [(t, "c_int64_t" if n == "flags" else k, n, ln) for t, k, n, ln in fm]) != []
    from wumingpic2d_b200.api import WM_NSP_MAX, WmConfig
    ctype = {"int32_t": C.c_int32, "double": C.c_double, "int64_t": C.c_int64}
    want = [(n, ctype[t] * WM_NSP_MAX if ln else ctype[t]) for t, n, ln in cm]
    assert [(n, t) for n, t in WmConfig._fields_] == want
    # the enumerators and flags carry the same values on all three sides
    hdr, f90 = open(B.HEADER).read(), open(B.CABI).read()
    import wumingpic2d_b200.api as api
    for name in ("WM_BC_PERIODIC", "WM_BC_RECONNECTION", "WM_BC_SHOCK", "WM_FLAG_EXACT_PUSH", "WM_NSP_MAX"):
        c_val = int(re.search(r"\b%s\b\s*=?\s*(\d+)" % name, hdr).group(1))
        f_val = int(re.search(r"\b%s\s*=\s*(\d+)" % name, f90).group(1))
        assert c_val == f_val == getattr(api, name), name
