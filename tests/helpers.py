"""Shared helpers for the parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import numpy as np

import oracle_lib as O


def flatten_by_id(up, np2):
    """(nsp, nyl, np, 6) host array + np2 -> ids, species, rec[n,5], row index; sorted by (isp, id)."""
    ids, sp, rec, row = [], [], [], []
    nsp, nyl = np2.shape
    for isp in range(nsp):
        for jl in range(nyl):
            n = int(np2[isp, jl])
            blk = up[isp, jl, :n, :]
            ids.append(blk[:, 5].copy().view(np.int64))
            sp.append(np.full(n, isp, np.int64))
            rec.append(blk[:, :5].copy())
            row.append(np.full(n, jl, np.int64))
    ids, sp, rec, row = map(np.concatenate, (ids, sp, rec, row))
    order = np.lexsort((ids, sp))
    return ids[order], sp[order], rec[order], row[order]


def oracle_state(w, rank=0):
    """Copies of one rank's host arrays."""
    return dict(up=w.array(rank, O.UP).copy(), gp=w.array(rank, O.GP).copy(), uf=w.array(rank, O.UF).copy(),
                np2=w.array(rank, O.NP2).copy(), cumcnt=w.array(rank, O.CUMCNT).copy(),
                uj=w.array(rank, O.UJ).copy(), df=w.array(rank, O.DF).copy())


def make_world(nx, ny, ppc, nranks=1, steps=0, seed=20260117, **kw):
    prm = O.weibel_params(nx, ny, ppc, nranks=nranks, **kw)
    w = O.World(prm)
    w.ic_weibel(seed)
    if steps:
        w.step(steps)
    return prm, w


def rel_to_max(a, b):
    """max |a-b| per last-axis component relative to that component's max |b| (norm-wise)."""
    d = np.abs(a - b).reshape(-1, a.shape[-1]).max(axis=0)
    s = np.abs(b).reshape(-1, b.shape[-1]).max(axis=0)
    return d / np.where(s > 0, s, 1.0)


def particle_err(rec_a, rec_b, xscale, vth):
    """position error relative to the domain scale, momentum error relative to max(|u|, vth)."""
    ex = np.abs(rec_a[:, :2] - rec_b[:, :2]).max() / xscale
    eu = (np.abs(rec_a[:, 2:] - rec_b[:, 2:]) / np.maximum(np.abs(rec_b[:, 2:]), vth)).max()
    return ex, eu


def make_wall_world(nx, ny, ppc, nranks=1, steps=0, seed=7, b0=0.02, **kw):
    """A reconnection-type world (proj/reconnection: reflecting/conducting x walls, periodic y) with a
    neutral uniform plasma between the walls and a Harris-like By(x) = b0 tanh((x - xc)/lam).
    The IC is written into gp and bucket-sorted by the oracle (the restart path, app.f90:351-352)."""
    prm = O.weibel_params(nx, ny, ppc, nranks=nranks, **kw)
    prm["bc"] = O.BC_RECONNECTION
    w = O.World(prm)
    rng = np.random.default_rng(seed)
    nxgs, nygs = prm["nxgs"], prm["nygs"]
    xlo, xhi = nxgs + 1, nxgs + nx - 3          # particles live in cells nxs+1 .. nxe-2 (nxe = nxgs+nx-1)
    npr = ppc * (xhi - xlo + 1)
    gid = 0
    for rk in range(nranks):
        nys, nye = w.bounds(rk)
        gp, np2 = w.array(rk, O.GP), w.array(rk, O.NP2)
        uf = w.array(rk, O.UF)
        ii = np.arange(uf.shape[1]) + (nxgs - 2)
        uf[:, :, 1] = b0 * np.tanh((ii[None, :] - (nxgs + nx / 2.0)) / 3.0)
        for jl in range(nye - nys + 1):
            # the same sequence for any rank count: seed by global row
            r = np.random.default_rng([seed, nys + jl])
            x = r.uniform(xlo, xhi + 1, npr)
            y = (nys + jl) + r.uniform(0, 1, npr)
            for isp in range(2):
                u = r.normal(0.0, prm["vte"], (npr, 3))
                gp[isp, jl, :npr, 0], gp[isp, jl, :npr, 1] = x, y
                gp[isp, jl, :npr, 2:5] = u
                gp[isp, jl, :npr, 5].view(np.int64)[:] = -(np.arange(npr) + 1 + (nys - nygs + jl) * npr)
                np2[isp, jl] = npr
    w.sort_bucket()
    for rk in range(nranks):
        w.array(rk, O.GP)[...] = w.array(rk, O.UP)
    if steps:
        w.step(steps)
    return prm, w


def make_shock_world(nx, ny, ppc, u0=-0.3, nranks=1, seed=11, b0=0.02, nxe=None, **kw):
    """A shock-type world (proj/shock: reflecting wall on the left, injection wall at xend on the right, periodic
    y): a neutral plasma that drifts toward the left wall with four-velocity u0 < 0 in a uniform By, Ez = v0 By
    (the upstream state of proj/shock/app.f90:400-470, without the driver's per-step injection of new
    particles).  The IC is written into gp and bucket-sorted by the oracle."""
    prm = O.weibel_params(nx, ny, ppc, nranks=nranks, **kw)
    prm["bc"] = O.BC_SHOCK
    prm["u0"] = u0
    w = O.World(prm)
    w.set_u_inject(u0)
    nxgs, nygs = prm["nxgs"], prm["nygs"]
    if nxe is None:
        nxe = nxgs + nx - 1
    else:                                       # the box of proj/shock starts short and grows (relocate)
        assert w.lib.orc_set_xrange(w.h, nxgs, nxe) == 0
    v0 = u0 / np.sqrt(1 + u0 * u0 / prm["c"] ** 2)
    xlo, xhi = nxgs + 1, nxe - 1                # particles live in cells nxs+1 .. nxe-1
    npr = ppc * (xhi - xlo + 1)
    for rk in range(nranks):
        nys, nye = w.bounds(rk)
        gp, np2 = w.array(rk, O.GP), w.array(rk, O.NP2)
        uf = w.array(rk, O.UF)
        uf[:, :, 1] = b0                          # By
        uf[:, :, 5] = -v0 * b0 / prm["c"]         # Ez = -v0 By / c   (proj/shock/app.f90: uf(6) = -v0*uf(2)/c)
        for jl in range(nye - nys + 1):
            r = np.random.default_rng([seed, nys + jl])
            x = r.uniform(xlo, xhi + 1, npr)
            y = (nys + jl) + r.uniform(0, 1, npr)
            for isp in range(2):
                u = r.normal(0.0, prm["vte"], (npr, 3))
                gam = np.sqrt(1 + (u ** 2).sum(axis=1) / prm["c"] ** 2)
                u[:, 0] = (u[:, 0] + v0 * gam) / np.sqrt(1 - v0 * v0 / prm["c"] ** 2)   # Lorentz boost, app.f90:452
                gp[isp, jl, :npr, 0], gp[isp, jl, :npr, 1] = x, y
                gp[isp, jl, :npr, 2:5] = u
                gp[isp, jl, :npr, 5].view(np.int64)[:] = -(np.arange(npr) + 1 + (nys - nygs + jl) * npr)
                np2[isp, jl] = npr
    w.sort_bucket()
    for rk in range(nranks):
        w.array(rk, O.GP)[...] = w.array(rk, O.UP)
    return prm, w


def shock_relocate(prm, up, np2, uf, nxe_new, n0, u0, b0, seed):
    """The state changes of `relocate` (proj/shock/app.f90:611-680) on host arrays of one rank holding all rows:
    n0 new particles per row and species in the cell nxe_new-1 (evenly spaced in x, drifting Maxwellian; ions and
    electrons at the same positions), upstream By / Ez in the new columns.  Deterministic stand-in for the
    driver's RNG; returns the appended records per species (for wm_append_particles)."""
    nxgs, nygs = prm["nxgs"], prm["nygs"]
    v0 = u0 / np.sqrt(1 + u0 * u0 / prm["c"] ** 2)
    nsp, nyl = np2.shape
    added = [[] for _ in range(nsp)]
    for jl in range(nyl):
        r = np.random.default_rng([seed, nxe_new, jl])
        x = (nxe_new - 1) + (np.arange(n0) + 0.5) / n0
        y = (nygs + jl) + r.uniform(0, 1, n0)
        for isp in range(nsp):
            u = r.normal(0.0, prm["vte"], (n0, 3))
            gam = np.sqrt(1 + (u ** 2).sum(axis=1) / prm["c"] ** 2)
            u[:, 0] = (u[:, 0] + v0 * gam) / np.sqrt(1 - v0 * v0 / prm["c"] ** 2)
            k = np2[isp, jl]
            up[isp, jl, k:k + n0, 0], up[isp, jl, k:k + n0, 1] = x, y
            up[isp, jl, k:k + n0, 2:5] = u
            up[isp, jl, k:k + n0, 5].view(np.int64)[:] = -(10 ** 7 * nxe_new + jl * n0 + np.arange(n0) + 1)
            np2[isp, jl] = k + n0
            added[isp].append(up[isp, jl, k:k + n0].copy())
    for i in (nxe_new - 1, nxe_new):              # uf(2,...) = By, uf(6,...) = Ez in the columns that became active
        uf[:, i - (nxgs - 2), 1] = b0
    uf[:, (nxe_new - 1) - (nxgs - 2), 5] = -v0 * b0 / prm["c"]
    return [np.concatenate(a) for a in added]


def make_langmuir_world(nx=32, ny=4, ppc=16, v0=1e-3, m=1, wpe=0.1):
    """Cold uniform plasma with heavy ions: both species on the same quiet 4 x 4 lattice per cell (ppc = 16), the
    electrons with ux = v0 sin(k x).  Written into gp and bucket-sorted by the oracle (the restart path)."""
    assert ppc == 16
    prm = O.weibel_params(nx, ny, ppc, mass_ratio=1e8, omega_pe=wpe, cap_factor=4.0)
    w = O.World(prm)
    nxgs, nygs = prm["nxgs"], prm["nygs"]
    k = 2 * np.pi * m / nx
    gp, np2 = w.array(0, O.GP), w.array(0, O.NP2)
    sub = (np.arange(4) + 0.5) / 4
    x = (nxgs + np.arange(nx)[:, None, None] + sub[None, :, None] + 0 * sub[None, None, :]).reshape(-1)
    yo = (0 * np.arange(nx)[:, None, None] + 0 * sub[None, :, None] + sub[None, None, :]).reshape(-1)
    npr = x.size
    for jl in range(ny):
        for isp in range(2):
            gp[isp, jl, :npr, 0] = x
            gp[isp, jl, :npr, 1] = nygs + jl + yo
            gp[isp, jl, :npr, 2:5] = 0.0
            if isp == 1:
                gp[isp, jl, :npr, 2] = v0 * np.sin(k * (x - nxgs))
            gp[isp, jl, :npr, 5].view(np.int64)[:] = -(np.arange(npr) + 1 + jl * npr)
            np2[isp, jl] = npr
    w.sort_bucket()
    w.array(0, O.GP)[...] = w.array(0, O.UP)
    return prm, w


def langmuir_fit(series, delt):
    """Frequency from the zero crossings of Ex(t) at the probe, and the peak |Ex| of every half period."""
    e = np.asarray(series)
    sgn = np.sign(e)
    idx = np.where(sgn[1:] * sgn[:-1] < 0)[0]
    tz = idx + e[idx] / (e[idx] - e[idx + 1])      # linear interpolation of the crossing times
    omega = np.pi / (np.diff(tz).mean() * delt)
    peaks = [np.abs(e[int(a):int(b) + 1]).max() for a, b in zip(tz[:-1], tz[1:])]
    return omega, peaks, len(idx)


def em_wave_in_plasma_setup(m, nx=32, wpe=0.1, e0=1e-3):
    """Cold plasma at rest (the Langmuir lattice with v0 = 0) + a standing wave Ez = e0 cos(k x), B = 0.  Returns
    (prm, world, uf with the wave, omega of the scheme, textbook omega).  Ez then oscillates at the frequency of a light
    wave in a plasma: textbook omega^2 = omega_pe^2 + c^2 k^2; on the grid k -> kappa = 2 sin(k delx / 2) / delx and the
    implicit time stepping turns s = omega dt into atan(s (1 - gfac)) + atan(s gfac) (exact for the vacuum part, see
    test_known_answer_vacuum_wave_dispersion; the plasma current enters the same way to 0.03 %)."""
    prm, w = make_langmuir_world(nx, 4, 16, v0=0.0, m=1, wpe=wpe)
    k = 2 * np.pi * m / nx
    uf = w.array(0, O.UF).copy()
    ii = np.arange(uf.shape[1]) + (prm["nxgs"] - 2)
    uf[...] = 0.0
    uf[:, :, 5] = e0 * np.cos(k * (ii + 0.5))[None, :]
    dt, c, g = prm["delt"], prm["c"], prm["gfac"]
    kap = 2 * np.sin(k * prm["delx"] / 2) / prm["delx"]
    s_ = dt * np.sqrt(wpe ** 2 + (c * kap) ** 2)
    om_scheme = (np.arctan(s_ * (1 - g)) + np.arctan(s_ * g)) / dt
    om_text = np.sqrt(wpe ** 2 + (c * k) ** 2)
    return prm, w, uf, om_scheme, om_text


def harris_params(nx, ny, nbg, ncs, nranks=1, mass_ratio=16.0, alpha=2.0, rtemp=0.2, lcs=0.5, cfl=0.5):
    """Physical set-up of proj/reconnection/app.f90:292-311 (constants c = delx = 1, gfac = 0.501, cfl = 0.5 of
    app.f90:64-68) as a parameter dict for the oracle / Context: adds vti, vte, b0, lcs_cells, n0 = nbg + ncs, and
    np = n0 * nx (app.f90:249)."""
    import math
    c, delx, gfac = 1.0, 1.0, 0.501
    r = [mass_ratio, 1.0]
    delt = cfl * delx / c
    vte = math.sqrt(rtemp) * c / (math.sqrt(1 + rtemp) * alpha)
    vti = vte * math.sqrt(r[1] / r[0]) / math.sqrt(rtemp)
    wpe = vte / delx / math.sqrt(2.0)
    wpi = wpe * math.sqrt(r[1] / r[0])
    wge = wpe / alpha
    wgi = wge / mass_ratio
    n0 = nbg + ncs
    q = [+math.sqrt(r[0] / (4 * math.pi * n0 / delx ** 2)) * wpi, -math.sqrt(r[1] / (4 * math.pi * n0 / delx ** 2)) * wpe]
    b0 = r[0] * c / q[0] * wgi
    return dict(nx=nx, ny=ny, nranks=nranks, n0=n0, np=n0 * nx, nsp=2, delx=delx, delt=delt, c=c, gfac=gfac, q=q, r=r, b0=b0,
                vti=vti, vte=vte, t_ani=1.0, nxgs=2, nygs=2, bc=O.BC_RECONNECTION, nbg=nbg, ncs=ncs, rtemp=rtemp,
                lcs_cells=lcs * c / wpi)


def shock_params(nx, ny, n0, nranks=1, u_inject=0.4, mass_ratio=1.0, sigma_e=0.1, omega_pe=0.1, v_the=0.05, v_thi=0.05,
                 theta_bn=90.0, phi_bn=90.0, l_damp_ini=10.0, cap_factor=5.0):
    """Physical set-up of proj/shock/app.f90:317-334 (c = delx = 1, cfl = 1, gfac = 0.501): u0 = -|u_inject|, v0 = u0 / gam0,
    charges scaled by gam0, b0 from sigma_e; np = n_ppc * nx * 5 (app.f90:277)."""
    import math
    c, delx, gfac, cfl = 1.0, 1.0, 0.501, 1.0
    delt = cfl * delx / c
    u0 = -abs(u_inject)
    gam0 = math.sqrt(1 + u0 * u0 / (c * c))
    v0 = u0 / gam0
    wpe = omega_pe
    wge = omega_pe * math.sqrt(sigma_e)
    wpi = wpe / math.sqrt(mass_ratio)
    wgi = wge / mass_ratio
    r = [mass_ratio, 1.0]
    q = [+math.sqrt(gam0 * r[0] / (4 * math.pi * n0 / delx ** 2)) * wpi, -math.sqrt(gam0 * r[1] / (4 * math.pi * n0 / delx ** 2)) * wpe]
    b0 = r[0] * c / q[0] * wgi * gam0
    return dict(nx=nx, ny=ny, nranks=nranks, n0=n0, np=int(n0 * nx * cap_factor), nsp=2, delx=delx, delt=delt, c=c, gfac=gfac,
                q=q, r=r, b0=b0, vti=v_thi, vte=v_the, t_ani=1.0, nxgs=2, nygs=2, bc=O.BC_SHOCK, u0=u0, v0=v0,
                theta=math.radians(theta_bn), phi=math.radians(phi_bn), l_damp=l_damp_ini)


def load_state_into_oracle(w, up, np2, cumcnt, uf, rank=0):
    """A sorted device state (download_particles + download_field) becomes the oracle's state (up, gp, np2, cumcnt, uf)."""
    for which, a in ((O.UP, up), (O.GP, up), (O.NP2, np2), (O.CUMCNT, cumcnt), (O.UF, uf)):
        dst = w.array(rank, which)
        assert dst.shape == a.shape, (which, dst.shape, a.shape)
        dst[...] = a
