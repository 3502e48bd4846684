"""Second, independent restatement of the reference hot path in vectorised numpy (TEST INFRASTRUCTURE).

Written from the Fortran sources, not from oracle/wm_oracle.cpp, so that the C++ oracle is checked
by something other than itself (the reference ships no golden vectors for this path and cannot be
built here).  One rank, periodic boundaries.  Arrays are C-ordered views of the reference's
Fortran arrays: uf[j, i, k] = uf(k+1, nxgs-2+i, nys-2+j) etc.

  push          common/particle.f90:69-161
  deposit       common/field.f90:215-310
  bc_curre      common/boundary_periodic.f90:357-508   (ring of one rank)
  field_fdtd_i  common/field.f90:121-184, cgm :319-461, bc_dfield boundary_periodic.f90:251-354,
                bc_phi :511-568
"""
import math

import numpy as np


def shape3(h):
    """particle.f90:97-105 / field.f90:224-236: S(-1), S(0), S(+1) at offset h."""
    return 0.5 * (0.5 - h) * (0.5 - h), 0.75 - h * h, 0.5 * (0.5 + h) * (0.5 + h)


def cell_centre_fields(uf):
    """particle.f90:69-81.  Returns tmp with the same padded indexing as uf (last row/col unused)."""
    t = np.zeros_like(uf)
    a = uf
    t[:-1, :-1, 0] = 0.5 * (+a[:-1, :-1, 0] + a[1:, :-1, 0])
    t[:-1, :-1, 1] = 0.5 * (+a[:-1, :-1, 1] + a[:-1, 1:, 1])
    t[:-1, :-1, 2] = 0.25 * (+a[:-1, :-1, 2] + a[:-1, 1:, 2] + a[1:, :-1, 2] + a[1:, 1:, 2])
    t[:-1, :-1, 3] = 0.5 * (+a[:-1, :-1, 3] + a[:-1, 1:, 3])
    t[:-1, :-1, 4] = 0.5 * (+a[:-1, :-1, 4] + a[1:, :-1, 4])
    t[:-1, :-1, 5] = a[:-1, :-1, 5]
    return t


def push(x, y, u, ci, cj, isp, uf, prm, delt=None):
    """particle.f90:83-169 for particles with sorted cell (ci, cj) (global indices) and species isp."""
    delt = prm["delt"] if delt is None else delt
    c = prm["c"]
    q = np.asarray(prm["q"])[isp]
    r = np.asarray(prm["r"])[isp]
    nxgs, nygs = prm["nxgs"], prm["nygs"]
    tmp = cell_centre_fields(uf)
    fac1 = q / r * 0.5 * delt
    txxx = fac1 * fac1
    fac2 = q * delt / r
    shx = shape3(x - 0.5 - ci)
    shy = shape3(y - 0.5 - cj)
    pi, pj = ci - (nxgs - 2), cj - (nygs - 2)  # padded indices
    f = []
    for k in range(6):
        acc = None
        for b in (-1, 0, 1):
            row = (+tmp[pj + b, pi - 1, k] * shx[0] + tmp[pj + b, pi, k] * shx[1] + tmp[pj + b, pi + 1, k] * shx[2]) * shy[b + 1]
            acc = row if acc is None else acc + row
        f.append(acc)
    bpx, bpy, bpz, epx, epy, epz = f
    uvm1 = u[:, 0] + fac1 * epx
    uvm2 = u[:, 1] + fac1 * epy
    uvm3 = u[:, 2] + fac1 * epz
    gam = np.sqrt(c * c + uvm1 * uvm1 + uvm2 * uvm2 + uvm3 * uvm3)
    igam = 1. / gam
    fac1r = fac1 * igam
    fac2r = fac2 / (gam + txxx * (bpx * bpx + bpy * bpy + bpz * bpz) * igam)
    uvm4 = uvm1 + fac1r * (+uvm2 * bpz - uvm3 * bpy)
    uvm5 = uvm2 + fac1r * (+uvm3 * bpx - uvm1 * bpz)
    uvm6 = uvm3 + fac1r * (+uvm1 * bpy - uvm2 * bpx)
    uvm1 = uvm1 + fac2r * (+uvm5 * bpz - uvm6 * bpy)
    uvm2 = uvm2 + fac2r * (+uvm6 * bpx - uvm4 * bpz)
    uvm3 = uvm3 + fac2r * (+uvm4 * bpy - uvm5 * bpx)
    un = np.stack([uvm1 + fac1 * epx, uvm2 + fac1 * epy, uvm3 + fac1 * epz], axis=1)
    g = 1. / np.sqrt(1.0 + (+un[:, 0] * un[:, 0] + un[:, 1] * un[:, 1] + un[:, 2] * un[:, 2]) / (c * c))
    return x + un[:, 0] * delt * g, y + un[:, 1] * delt * g, un


def _ds(xn, c0, s0):
    """field.f90:238-251,268: shifted new shape minus old shape, as 5 arrays (-2..2)."""
    i2 = np.trunc(xn).astype(np.int64)
    s1, s2, s3 = shape3(xn - 0.5 - i2)
    inc = i2 - c0
    m1 = -(inc - np.abs(inc)) * 0.5
    m2 = -np.abs(inc) + 1.0
    m3 = (inc + np.abs(inc)) * 0.5
    d = [s1 * m1, s1 * m2 + s2 * m1, s2 * m2 + s3 * m1 + s1 * m3, s3 * m2 + s2 * m3, s3 * m3]
    z = np.zeros_like(xn)
    s0f = [z, s0[0], s0[1], s0[2], z]
    return [d[k] - s0f[k] for k in range(5)], s0f


def deposit(x, y, xn, yn, un, ci, cj, isp, prm, nx, nyl):
    """field.f90:215-310 -> uj[(nyl+4), (nx+4), 3] before any boundary treatment."""
    c = prm["c"]
    q = np.asarray(prm["q"])[isp]
    nxgs, nygs = prm["nxgs"], prm["nygs"]
    dsx, s0x = _ds(xn, ci, shape3(x - 0.5 - ci))
    dsy, s0y = _ds(yn, cj, shape3(y - 0.5 - cj))
    gvz = un[:, 2] / np.sqrt(1. + (+un[:, 0] * un[:, 0] + un[:, 1] * un[:, 1] + un[:, 2] * un[:, 2]) / (c * c))
    qf = q * prm["delx"] * (1.0 / prm["delt"])
    uj = np.zeros((nyl + 4, nx + 4, 3))
    pi, pj = ci - (nxgs - 2), cj - (nygs - 2)
    for jp in range(5):
        run = np.zeros_like(x)
        for ip in range(4):  # pjtmp(ip+1,jp) = pjtmp(ip,jp) - q*delx*d_delt*ds(ip,1)*(s0(jp,2)+0.5*ds(jp,2))
            run = run - qf * dsx[ip] * (s0y[jp] + 0.5 * dsy[jp])
            np.add.at(uj[:, :, 0], (pj + jp - 2, pi + ip + 1 - 2), run)
    for ip in range(5):
        run = np.zeros_like(x)
        for jp in range(4):
            run = run - qf * dsy[jp] * (s0x[ip] + 0.5 * dsx[ip])
            np.add.at(uj[:, :, 1], (pj + jp + 1 - 2, pi + ip - 2), run)
    fac = 1.0 / 3.0
    for jp in range(5):
        for ip in range(5):
            v = q * gvz * (+s0x[ip] * s0y[jp] + 0.5 * dsx[ip] * s0y[jp] + 0.5 * s0x[ip] * dsy[jp] + fac * dsx[ip] * dsy[jp])
            np.add.at(uj[:, :, 2], (pj + jp - 2, pi + ip - 2), v)
    return uj


def bc_curre(uj):
    """boundary_periodic.f90:357-508 with nup = ndown = this rank.  In place."""
    n = uj.shape[0] - 4  # rows nys..nye are 2..n+1
    lo, hi = uj[0:2].copy(), uj[n + 2:n + 4].copy()
    uj[n:n + 2] += lo          # ghosts nys-2,nys-1 -> (ndown's) nye-1,nye
    uj[2:4] += hi              # ghosts nye+1,nye+2 -> (nup's) nys,nys+1
    uj[n + 2:n + 4] = uj[2:4]  # refresh
    uj[0:2] = uj[n:n + 2]
    m = uj.shape[1] - 4
    uj[:, m] += uj[:, 0]
    uj[:, m + 1] += uj[:, 1]
    uj[:, 2] += uj[:, m + 2]
    uj[:, 3] += uj[:, m + 3]
    uj[:, 0] = uj[:, m]
    uj[:, 1] = uj[:, m + 1]
    uj[:, m + 2] = uj[:, 2]
    uj[:, m + 3] = uj[:, 3]
    return uj


def bc_dfield(df):
    """boundary_periodic.f90:251-354 with one rank.  In place."""
    n, m = df.shape[0] - 4, df.shape[1] - 4
    df[n + 2:n + 4] = df[2:4]
    df[0:2] = df[n:n + 2]
    df[:, 0:2] = df[:, m:m + 2]
    df[:, m + 2:m + 4] = df[:, 2:4]
    return df


def _bc_phi(a):
    """boundary_periodic.f90:511-568 on an array with a 1-cell halo, one rank."""
    a[-1, 1:-1] = a[1, 1:-1]
    a[0, 1:-1] = a[-2, 1:-1]
    a[:, 0] = a[:, -2]
    a[:, -1] = a[:, 1]


def cgm(df, gkl, f4, f5):
    """field.f90:347-459, literally.  df[(nyl+4),(nx+4),6] in/out, gkl[nyl,nx,3].  Returns iteration counts."""
    nyl, nx = gkl.shape[:2]
    ites = []
    for l in range(3):
        phi = np.zeros((nyl + 2, nx + 2))
        p = np.zeros((nyl + 2, nx + 2))
        phi[1:-1, 1:-1] = df[2:-2, 2:-2, l]
        b = f5 * gkl[:, :, l]
        sum_g = float(np.sum(b * b))
        eps = math.sqrt(sum_g) * 1e-6
        ite = 0
        _bc_phi(phi)
        I = (slice(1, -1), slice(1, -1))
        r = b + phi[:-2, 1:-1] + phi[1:-1, :-2] - f4 * phi[I] + phi[1:-1, 2:] + phi[2:, 1:-1]
        p[I] = r
        sumr_g = float(np.sum(r * r))
        if math.sqrt(sumr_g) > eps:
            while sum_g > eps:
                ite += 1
                _bc_phi(p)
                ap = -p[:-2, 1:-1] - p[1:-1, :-2] + f4 * p[I] - p[1:-1, 2:] - p[2:, 1:-1]
                sumr_g = float(np.sum(r * r))
                sum2_g = float(np.sum(p[I] * ap))
                av = sumr_g / sum2_g
                phi[I] = phi[I] + av * p[I]
                r = r - av * ap
                sum_g = math.sqrt(sumr_g)
                if ite >= 100:
                    raise RuntimeError("stop at cgm after ite_max")
                sum1_g = float(np.sum(r * r))
                bv = sum1_g / sumr_g
                p[I] = r + bv * p[I]
        df[2:-2, 2:-2, l] = phi[I]
        ites.append(ite)
    return ites


def field_fdtd_i(uf, uj, df, prm):
    """field.f90:121-184 after ele_cur; uj already has its boundary applied.  uf, df in/out."""
    c, delt, delx, gfac = prm["c"], prm["delt"], prm["delx"], prm["gfac"]
    pi = 4.0 * math.atan(1.0)
    f1 = c * delt / delx
    f2 = gfac * f1 * f1
    f3 = 4.0 * pi * delx / c
    f4 = 4.0 + (delx / (c * delt * gfac)) ** 2
    f5 = (delx / (c * delt * gfac)) ** 2
    C = (slice(2, -2), slice(2, -2))
    S = (slice(1, -3), slice(2, -2))   # j-1
    N = (slice(3, -1), slice(2, -2))   # j+1
    W = (slice(2, -2), slice(1, -3))   # i-1
    E = (slice(2, -2), slice(3, -1))   # i+1

    def lap(k):
        return +uf[S + (k,)] + uf[W + (k,)] - 4. * uf[C + (k,)] + uf[E + (k,)] + uf[N + (k,)]

    gkl = np.zeros(uf[C].shape[:2] + (3,))
    gkl[:, :, 0] = +f2 * (lap(0) + f3 * (-uj[S + (2,)] + uj[C + (2,)])) - f1 * (-uf[S + (5,)] + uf[C + (5,)])
    gkl[:, :, 1] = +f2 * (lap(1) - f3 * (-uj[W + (2,)] + uj[C + (2,)])) + f1 * (-uf[W + (5,)] + uf[C + (5,)])
    gkl[:, :, 2] = (+f2 * (lap(2) + f3 * (-uj[W + (1,)] + uj[C + (1,)] + uj[S + (0,)] - uj[C + (0,)]))
                    - f1 * (-uf[W + (4,)] + uf[C + (4,)] + uf[S + (3,)] - uf[C + (3,)]))
    ites = cgm(df, gkl, f4, f5)
    bc_dfield(df)
    df[C + (3,)] = +f1 * (+gfac * (-df[C + (2,)] + df[N + (2,)]) + (-uf[C + (2,)] + uf[N + (2,)])) - 4. * pi * delt * uj[C + (0,)]
    df[C + (4,)] = -f1 * (+gfac * (-df[C + (2,)] + df[E + (2,)]) + (-uf[C + (2,)] + uf[E + (2,)])) - 4. * pi * delt * uj[C + (1,)]
    df[C + (5,)] = (+f1 * (+gfac * (-df[C + (1,)] + df[E + (1,)] + df[C + (0,)] - df[N + (0,)])
                           + (-uf[C + (1,)] + uf[E + (1,)] + uf[C + (0,)] - uf[N + (0,)])) - 4. * pi * delt * uj[C + (2,)])
    bc_dfield(df)
    uf += df
    return ites


def flatten_sorted(up, np2, cumcnt, nxgs, nygs):
    """Host arrays (nsp, nyl, np, 6) + np2 + cumcnt -> flat per-particle arrays with their sorted cell."""
    xs, ys, us, ids, ci, cj, sp = [], [], [], [], [], [], []
    nsp, nyl = np2.shape
    for isp in range(nsp):
        for jl in range(nyl):
            n = int(np2[isp, jl])
            blk = up[isp, jl, :n]
            xs.append(blk[:, 0]); ys.append(blk[:, 1]); us.append(blk[:, 2:5]); ids.append(blk[:, 5].copy().view(np.int64))
            ci.append(np.repeat(np.arange(cumcnt.shape[2] - 1) + nxgs, np.diff(cumcnt[isp, jl])))
            cj.append(np.full(n, jl + nygs, np.int64))
            sp.append(np.full(n, isp, np.int64))
    cat = np.concatenate
    return cat(xs), cat(ys), cat(us), cat(ids), cat(ci).astype(np.int64), cat(cj), cat(sp)
