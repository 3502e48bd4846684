"""Device-side initial conditions and particle sources of the applications (BASELINE configs 3
and 4): the Harris current sheet of proj/reconnection/app.f90:368-456 and the shock driver's initial load / inject /
relocate of proj/shock/app.f90:406-470,611-850.  Two kinds of checks: (1) the generated state has the reference's
distributions (row counts, density profile, drifts, field profile); (2) the hot path on these strongly non-uniform loads
agrees with the oracle step by step (per-cell counts bit-exact, particles by ID / fields within tolerance).  -m gpu."""
import math

import numpy as np
import pytest

import oracle_lib as O
from helpers import flatten_by_id, harris_params, load_state_into_oracle, particle_err, rel_to_max, shock_params

pytestmark = pytest.mark.gpu


def _compare(c, w, prm, tol, na=None):
    assert c.cg_iters() == w.cg_iters(), (c.cg_iters(), w.cg_iters())
    up, np2, cum = c.download_particles()
    assert np.array_equal(np2, w.array(0, O.NP2))
    if na is None:
        assert np.array_equal(cum, w.array(0, O.CUMCNT)), "per-cell counts must be bit-exact"
    else:
        assert np.array_equal(cum[:, :, :na + 1], w.array(0, O.CUMCNT)[:, :, :na + 1]), "per-cell counts must be bit-exact"
    a, b = flatten_by_id(up, np2), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
    assert np.array_equal(a[0], b[0])
    ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
    assert ex <= tol and eu <= tol, (ex, eu)
    assert rel_to_max(c.download_field(), w.array(0, O.UF)).max() <= tol


def test_harris_sheet_on_device():
    import wumingpic2d_b200 as wm
    nx, ny, nbg, ncs = 64, 24, 8, 40
    prm = harris_params(nx, ny, nbg, ncs, lcs=0.25)
    lcs = prm["lcs_cells"]
    c = wm.Context.from_params(prm, bc=wm.WM_BC_RECONNECTION)
    c.ic_harris(11, nbg, ncs, lcs, prm["vti"], prm["vte"], prm["b0"], prm["rtemp"], 0.12)
    up, np2, cum = c.download_particles()
    uf = c.download_field()
    # ---- the distributions of app.f90:313,414-448
    npr = nbg * (nx - 1) + int(ncs * 2 * lcs)
    assert np.all(np2 == npr)
    nxgs = prm["nxgs"]
    x0 = 0.5 * (nxgs + nx - 1 + nxgs)
    xs = np.concatenate([up[0, j, :npr, 0] for j in range(ny)])
    assert xs.min() >= nxgs + 1 and xs.max() <= nxgs + nx - 2            # between the walls nxs+1 .. nxe-1
    ids, sp, rec, _ = flatten_by_id(up, np2)                             # (the order inside a cell is free, so match by id)
    assert np.array_equal(ids[sp == 0], ids[sp == 1])
    assert np.array_equal(rec[sp == 0][:, :2], rec[sp == 1][:, :2])      # ions and electrons at the same positions
    # density: background nbg (nx - 3) uniform + sheet with sech^2((x - x0) / lcs) / (2 lcs) over +- half the box
    edges = np.arange(nxgs + 1, nxgs + nx - 1)
    hist = np.histogram(xs, bins=edges)[0] / ny
    xc = 0.5 * (edges[:-1] + edges[1:])
    nsheet = npr - nbg * (nx - 3)
    t = math.tanh(0.5 * (nx - 3) / lcs)
    model = nbg + nsheet * (np.tanh((edges[1:] - x0) / lcs) - np.tanh((edges[:-1] - x0) / lcs)) / (2 * t)
    assert np.abs(hist - model).max() <= 5 * np.sqrt(model.max() / ny) + 1
    assert hist[np.argmin(np.abs(xc - x0))] > 3 * nbg                   # the sheet is there
    # drift: <uz> of the ions in the sheet follows f1 jz / density > 0, the electrons' the opposite sign
    core = np.abs(up[0, :, :npr, 0] - x0) < 0.5 * lcs
    assert up[0, :, :npr, 4][core].mean() > 0 > up[1, :, :npr, 4][core].mean()
    sdi = prm["vti"] / np.float32(np.sqrt(np.float32(2.0)))
    far = np.abs(up[0, :, :npr, 0] - x0) > 4 * lcs
    assert abs(up[0, :, :npr, 2][far].std() / sdi - 1) < 0.05
    # field: By = b0 tanh((x - x0) / lcs) + perturbation <= e1 b0, Ex = Ey = Ez = Bz = 0
    ii = np.arange(uf.shape[1]) + (nxgs - 2)
    assert np.abs(uf[:, :, 1] - prm["b0"] * np.tanh((ii - x0) / lcs)[None, :]).max() <= 0.13 * prm["b0"]
    assert np.all(uf[:, :, 2:] == 0)
    # ---- the hot path on this load against the oracle
    w = O.World(prm)
    load_state_into_oracle(w, up, np2, cum, uf)
    for it in range(4):
        w.step(1)
        c.step(1)
        _compare(c, w, prm, 1e-12 if it == 0 else 1e-10)
    g0, g1 = c.energy(), w.energy()
    assert np.allclose(g0, g1, rtol=1e-10)
    c.close(); w.close()


def test_shock_sources_on_device():
    import wumingpic2d_b200 as wm
    nx, ny, n0 = 48, 12, 6
    prm = shock_params(nx, ny, n0, u_inject=0.5, l_damp_ini=8.0)
    nxgs = prm["nxgs"]
    nxe = nxgs + 30
    c = wm.Context.from_params(prm, bc=wm.WM_BC_SHOCK, capacity=n0 * nx * ny * 6)
    c.ic_shock(5, n0, nxe, prm["v0"], prm["vti"], prm["vte"], prm["b0"], prm["theta"], prm["phi"], prm["l_damp"])
    c.set_u_inject(prm["u0"])
    assert c.xrange() == (nxgs, nxe)
    up, np2, cum = c.download_particles()
    uf = c.download_field()
    npr = n0 * (nxe - nxgs - 1)                                           # app.f90:330
    assert np.all(np2 == npr)
    x = np.sort(up[0, 0, :npr, 0])
    # evenly spaced, n0 per cell in nxs+1 .. nxe-1 as the reference's cumcnt says (app.f90:341-344; see gen_kernels.cu)
    assert np.allclose(x, nxgs + 1 + (nxe - nxgs - 1) * (np.arange(npr) + 0.5) / npr, rtol=0, atol=1e-12)
    assert np.all(np.diff(cum[0, 0])[1:nxe - nxgs] == n0) and cum[0, 0, 1] == 0
    gam0 = 1 / math.sqrt(1 - prm["v0"] ** 2)
    up_stream = up[0, :, :npr, 0] > nxgs + 2 * prm["l_damp"]
    assert abs(up[0, :, :npr, 2][up_stream].mean() / (gam0 * prm["v0"]) - 1) < 0.05     # boosted to u0 far upstream
    assert abs(up[0, :, :npr, 2][~up_stream & (up[0, :, :npr, 0] < nxgs + 0.5 * prm["l_damp"])].mean()) < 0.3 * abs(prm["u0"])
    by, bz = prm["b0"] * math.sin(prm["theta"]) * math.cos(prm["phi"]), prm["b0"] * math.sin(prm["theta"]) * math.sin(prm["phi"])
    assert np.allclose(uf[:, :, 2], bz) and np.allclose(uf[:, :, 1], by, atol=1e-18)
    w = O.World(prm)
    w.set_u_inject(prm["u0"])
    assert w.lib.orc_set_xrange(w.h, nxgs, nxe) == 0
    load_state_into_oracle(w, up, np2, cum, uf)
    pflux = n0 * abs(prm["v0"]) * prm["delt"] * ny
    for it in range(1, 7):
        w.step(1)
        c.step(1)
        _compare(c, w, prm, 1e-12 if it == 1 else 1e-9, na=nxe - nxgs + 1)
        n_before = np.array(c.particle_counts())
        c.shock_inject(5, it)                                             # proj/shock/app.f90:119-125
        n_inj = np.array(c.particle_counts()) - n_before
        assert n_inj[0] == n_inj[1] and int(pflux) <= n_inj[0] <= int(pflux) + 1
        if it % 2 == 0:
            c.shock_relocate(5, it)
            nxe += 1
            assert c.xrange() == (nxgs, nxe)
            assert np.all(np.array(c.particle_counts()) - n_before - n_inj == n0 * ny)
        up, np2, cum = c.download_particles()
        uf = c.download_field()
        ids0, ids1 = up[0, :, :, 5].view(np.int64), up[1, :, :, 5].view(np.int64)
        live0 = np.arange(up.shape[2])[None, :] < np2[0][:, None]
        assert (ids0[live0] < -(1 << 35)).sum() >= n_inj[0]              # the sources' ids carry the step in their high bits
        assert w.lib.orc_set_xrange(w.h, nxgs, nxe) == 0
        load_state_into_oracle(w, up, np2, cum, uf)
    assert c.rebuilds() >= 0
    c.close(); w.close()
