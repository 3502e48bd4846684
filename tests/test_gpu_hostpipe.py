"""wm_host_step, pipelined: the rows travel in chunks and the upload of the next chunks, the particle pass over the chunk
that has arrived and the download of the finished rows run side by side (wm_api.cu "pipelined host step").  It must give
what upload + wm_step + download gives -- checked against the oracle (common/particle.f90:83-177, common/field.f90:121-184,
common/sort.f90:36-82, common/boundary_periodic.f90:99-248) for every chunking of the slab: one chunk, chunks of one and two
tile rows, a ragged last tile row, slabs of one to three tile rows (where everything waits for the ring), the wall kinds,
and a two-slab ring through the loopback transport.  -m gpu."""
import threading

import numpy as np
import pytest

import oracle_lib as O
from helpers import flatten_by_id, make_shock_world, make_wall_world, make_world, oracle_state, particle_err, rel_to_max

pytestmark = pytest.mark.gpu


def _check_against(w, prm, up, uf, np2, cum, rank=0, tol=1e-11):
    assert np.array_equal(np2, w.array(rank, O.NP2)), "np2 must be bit-exact"
    assert np.array_equal(cum, w.array(rank, O.CUMCNT)), "cumcnt must be bit-exact"
    assert rel_to_max(uf, w.array(rank, O.UF)).max() <= tol
    a, b = flatten_by_id(up, np2), flatten_by_id(w.array(rank, O.UP), w.array(rank, O.NP2))
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])     # ids, and the row/species each one lives in
    ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
    assert ex <= tol and eu <= tol, (ex, eu)
    # every row of the host array is sorted by cell (what cumcnt promises the next particle__solv)
    for isp in range(up.shape[0]):
        for jl in range(up.shape[1]):
            n = np2[isp, jl]
            xs = up[isp, jl, :n, 0].astype(np.int64)
            assert np.all(np.diff(xs) >= 0)


@pytest.mark.parametrize("ny,rows,chunks", [(16, 16, 1), (40, 8, 5), (40, 16, 3), (44, 8, 6), (8, 8, 1), (20, 8, 3), (72, 24, 3)])
def test_pipelined_host_step_matches_oracle(ny, rows, chunks, monkeypatch):
    import wumingpic2d_b200 as wm
    monkeypatch.setenv("WM_HOSTPIPE_ROWS", str(rows))
    prm, w = make_world(40, ny, 10)    # (from the IC: the CG warm start df is a SAVE variable that no host array carries)
    s = oracle_state(w)
    c = wm.Context.from_params(prm)
    up, uf, np2, cum = s["up"].copy(), s["uf"].copy(), s["np2"].copy(), s["cumcnt"].copy()
    for it in range(3):
        w.step(1)
        c.host_step(up, uf, np2, cum)
        assert c.host_pipe_chunks() == chunks
        _check_against(w, prm, up, uf, np2, cum, tol=1e-12 if it == 0 else 1e-10)
    # the device state after a pipelined step is the resident state: plain steps carry on from it
    w.step(2)
    c.step(2)
    upd, np2d, cumd = c.download_particles()
    _check_against(w, prm, upd, c.download_field(), np2d, cumd, tol=1e-10)
    c.close()
    w.close()


def test_pipelined_equals_sequential(monkeypatch):
    """the same call with WM_HOSTPIPE=0 (upload, wm_step, download): identical index arrays, the same particles"""
    import wumingpic2d_b200 as wm
    prm, w = make_world(56, 44, 12)
    s = oracle_state(w)
    out = []
    for pipe in ("1", "0"):
        monkeypatch.setenv("WM_HOSTPIPE", pipe)
        monkeypatch.setenv("WM_HOSTPIPE_ROWS", "8")
        c = wm.Context.from_params(prm)
        up, uf, np2, cum = s["up"].copy(), s["uf"].copy(), s["np2"].copy(), s["cumcnt"].copy()
        c.host_step(up, uf, np2, cum)     # (one step: a second one starts from fields that differ in the last bit, the current
        #                                    is summed with floating-point atomics)
        assert (c.host_pipe_chunks() > 0) == (pipe == "1")
        out.append((up, uf, np2, cum))
        c.close()
    (u1, f1, n1, c1), (u0, f0, n0, c0) = out
    assert np.array_equal(n1, n0) and np.array_equal(c1, c0)
    a, b = flatten_by_id(u1, n1), flatten_by_id(u0, n0)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
    assert np.array_equal(a[2], b[2]), "same kernels, same arithmetic per particle: the records are bit-identical"
    assert rel_to_max(f1, f0).max() <= 1e-12        # (the current is summed with floating-point atomics in both)
    w.close()


@pytest.mark.parametrize("kind", ["reconnection", "shock"])
def test_pipelined_host_step_walls(kind, monkeypatch):
    import wumingpic2d_b200 as wm
    monkeypatch.setenv("WM_HOSTPIPE_ROWS", "8")
    if kind == "reconnection":
        prm, w = make_wall_world(40, 28, 10)
        c = wm.Context.from_params(prm)
    else:
        prm, w = make_shock_world(40, 28, 10, u0=-0.3)
        c = wm.Context.from_params(prm, bc=wm.WM_BC_SHOCK)
        c.set_u_inject(-0.3)
    s = oracle_state(w)
    up, uf, np2, cum = s["up"].copy(), s["uf"].copy(), s["np2"].copy(), s["cumcnt"].copy()
    for it in range(4):
        w.step(1)
        c.host_step(up, uf, np2, cum)
        assert c.host_pipe_chunks() == 4
        _check_against(w, prm, up, uf, np2, cum, tol=1e-12 if it == 0 else 1e-10)
    c.close()
    w.close()


def test_pipelined_host_step_bad_input_is_an_error(monkeypatch):
    """cumcnt that does not describe the rows is refused (as wm_upload_particles_sorted refuses it), and the context says so
    on the next resident call instead of stepping a half-loaded state"""
    import wumingpic2d_b200 as wm
    monkeypatch.setenv("WM_HOSTPIPE_ROWS", "8")
    prm, w = make_world(40, 24, 8)
    s = oracle_state(w)
    c = wm.Context.from_params(prm)
    up, uf, np2, cum = s["up"].copy(), s["uf"].copy(), s["np2"].copy(), s["cumcnt"].copy()
    cum[0, 13, 5:9] += 1
    with pytest.raises(wm.WmError):
        c.host_step(up, uf, np2, cum)
    with pytest.raises(wm.WmError):
        c.step(1)
    # and a good call afterwards works
    up, uf, np2, cum = s["up"].copy(), s["uf"].copy(), s["np2"].copy(), s["cumcnt"].copy()
    w.step(1)
    c.host_step(up, uf, np2, cum)
    _check_against(w, prm, up, uf, np2, cum, tol=1e-12)
    c.close()
    w.close()


def test_pipelined_host_step_ring_loopback(monkeypatch):
    """two slabs of one process: the leavers of the edge rows go through the ring exchange at the end of the pipelined pass"""
    import wumingpic2d_b200 as wm
    monkeypatch.setenv("WM_HOSTPIPE_ROWS", "8")
    n = 2
    prm = O.weibel_params(40, 8 * 5 + 3, 10, nranks=n)
    w = O.World(prm)
    w.ic_weibel(20260117)
    init = [dict(up=w.array(r, O.UP).copy(), np2=w.array(r, O.NP2).copy(), cum=w.array(r, O.CUMCNT).copy(), uf=w.array(r, O.UF).copy())
            for r in range(n)]
    nsteps = 3
    ref = []
    for it in range(nsteps):
        w.step(1)
        ref.append([dict(up=w.array(r, O.UP).copy(), np2=w.array(r, O.NP2).copy(), cum=w.array(r, O.CUMCNT).copy(),
                         uf=w.array(r, O.UF).copy()) for r in range(n)])
    group = wm.LoopbackGroup(n)
    errs = [None] * n

    def rank_main(r):
        try:
            nys, nye = w.bounds(r)
            c = wm.Context.from_params(prm, nys=nys, nye=nye, nrank=r, nsize=n, device=0)
            c.comm_init_loopback(group)
            up, uf, np2, cum = init[r]["up"], init[r]["uf"], init[r]["np2"], init[r]["cum"]
            for it in range(nsteps):
                c.host_step(up, uf, np2, cum)
                assert c.host_pipe_chunks() >= 3
                q = ref[it][r]
                assert np.array_equal(np2, q["np2"]) and np.array_equal(cum, q["cum"]), "counts differ on rank %d step %d" % (r, it)
                a, b = flatten_by_id(up, np2), flatten_by_id(q["up"], q["np2"])
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
                ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
                tol = 1e-12 if it == 0 else 1e-10
                assert ex <= tol and eu <= tol, (r, it, ex, eu)
                assert rel_to_max(uf, q["uf"]).max() <= tol
            c.close()
        except BaseException as ex:  # noqa: BLE001 -- reported by the main thread
            errs[r] = ex

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(n)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=600)
    group.close()
    for r, e in enumerate(errs):
        if e is not None:
            raise AssertionError("rank %d: %r" % (r, e)) from e
    assert all(not t.is_alive() for t in th)
    w.close()


def test_host_register_pageable_arrays(monkeypatch):
    """wm_host_register page-locks the caller's (numpy = pageable) arrays: same results, registering twice is not an error"""
    import wumingpic2d_b200 as wm
    from wumingpic2d_b200.api import host_register, host_unregister
    monkeypatch.setenv("WM_HOSTPIPE_ROWS", "8")
    prm, w = make_world(40, 24, 8)
    s = oracle_state(w)
    c = wm.Context.from_params(prm)
    up, uf, np2, cum = s["up"].copy(), s["uf"].copy(), s["np2"].copy(), s["cumcnt"].copy()
    for a in (up, uf, np2, cum, up):
        host_register(a)
    for it in range(2):
        w.step(1)
        c.host_step(up, uf, np2, cum)
        _check_against(w, prm, up, uf, np2, cum, tol=1e-12 if it == 0 else 1e-10)
    for a in (up, uf, np2, cum):
        host_unregister(a)
    c.close()
    w.close()


def test_pipelined_host_step_with_overflowing_segments(monkeypatch):
    """hardly any segment slack: segments overflow inside a pipelined step, the parked records are not in the rows that have
    already gone back -> the layout is rebuilt and everything is fetched again (correct, slow)"""
    import wumingpic2d_b200 as wm
    monkeypatch.setenv("WM_HOSTPIPE_ROWS", "8")
    monkeypatch.setenv("WM_SLACK", "0.01")
    prm, w = make_world(40, 40, 17)
    s = oracle_state(w)
    c = wm.Context.from_params(prm)
    up, uf, np2, cum = s["up"].copy(), s["uf"].copy(), s["np2"].copy(), s["cumcnt"].copy()
    for it in range(6):
        w.step(1)
        c.host_step(up, uf, np2, cum)
        assert c.host_pipe_chunks() == 5
        _check_against(w, prm, up, uf, np2, cum, tol=1e-12 if it == 0 else 1e-9)
    assert c.rebuilds() > 0, "the overflow branch was not exercised"
    c.close()
    w.close()
