"""ctypes binding of the CPU oracle (oracle/wm_oracle.h).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs import this module.  The product package never does.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

NSP_MAX = 4
UP, GP, UF, DF, UJ, GKL, MOM = range(7)
NP2, CUMCNT = 16, 17
BC_PERIODIC, BC_RECONNECTION, BC_SHOCK = 0, 1, 2


class OrcConfig(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nxgs", C.c_int32), ("nygs", C.c_int32),
        ("nranks", C.c_int32), ("np", C.c_int32), ("nsp", C.c_int32), ("bc", C.c_int32),
        ("delx", C.c_double), ("delt", C.c_double), ("c", C.c_double), ("gfac", C.c_double),
        ("q", C.c_double * NSP_MAX), ("r", C.c_double * NSP_MAX),
    ]


def build(force=False):
    """Compile the oracle with the recipe in oracle/Makefile (g++, seconds)."""
    out = os.path.join(ORACLE_DIR, "_build", "libwm_oracle.so")
    if force or not os.path.exists(out) or not os.path.exists(out.replace(".so", "_fast.so")):
        subprocess.check_call(["make", "-C", ORACLE_DIR], stdout=subprocess.DEVNULL)
    return out


_libs = {}


def load(fast=False):
    key = "fast" if fast else "parity"
    if key in _libs:
        return _libs[key]
    path = build()
    if fast:
        path = path.replace(".so", "_fast.so")
    lib = C.CDLL(path)
    P = C.c_void_p
    lib.orc_create.restype = P
    lib.orc_create.argtypes = [C.POINTER(OrcConfig)]
    lib.orc_destroy.argtypes = [P]
    lib.orc_bounds.argtypes = [P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.orc_array.restype = P
    lib.orc_array.argtypes = [P, C.c_int, C.c_int, C.POINTER(C.c_int64)]
    for name in ("orc_particle_solv", "orc_ele_cur", "orc_bc_curre", "orc_bc_particle_x",
                 "orc_sort_bucket", "orc_mom_accl", "orc_mom_nvt", "orc_bc_mom"):
        getattr(lib, name).argtypes = [P]
        getattr(lib, name).restype = None
    for name in ("orc_field_fdtd_i", "orc_bc_particle_y"):
        getattr(lib, name).argtypes = [P]
        getattr(lib, name).restype = C.c_int
    lib.orc_bc_injection.argtypes = [P, C.c_double]
    lib.orc_bc_injection.restype = None
    lib.orc_set_u_inject.argtypes = [P, C.c_double]
    lib.orc_set_u_inject.restype = None
    lib.orc_set_xrange.argtypes = [P, C.c_int, C.c_int]
    lib.orc_set_xrange.restype = C.c_int
    lib.orc_step.argtypes = [P, C.c_int]
    lib.orc_step.restype = C.c_int
    lib.orc_stage_times.argtypes = [P, C.POINTER(C.c_double), C.c_int]
    lib.orc_cg_iters.argtypes = [P, C.POINTER(C.c_int32)]
    lib.orc_energy.argtypes = [P, C.POINTER(C.c_double)]
    lib.orc_gauss_residual.argtypes = [P, C.POINTER(C.c_double)]
    lib.orc_gauss_residual.restype = C.c_double
    lib.orc_ic_weibel.argtypes = [P, C.c_uint64, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]
    lib.orc_rng_hash.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_int]
    lib.orc_rng_hash.restype = C.c_uint64
    lib.orc_rng_uniform.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_int]
    lib.orc_rng_uniform.restype = C.c_double
    lib.orc_num_threads.restype = C.c_int
    lib.orc_set_num_threads.argtypes = [C.c_int]
    _libs[key] = lib
    return lib


def weibel_params(nx, ny, n_ppc, nranks=1, mass_ratio=1.0, sigma_e=0.0, omega_pe=0.1,
                  v_the=0.1, v_thi=0.1, t_ani=5.0, cap_factor=5.0, c=1.0, gfac=0.501,
                  cfl=1.0, delx=1.0):
    """Physical set-up of proj/weibel/app.f90:245-255,292-303 as a dict."""
    delt = cfl * delx / c
    wpe = omega_pe
    wge = omega_pe * math.sqrt(sigma_e)
    wpi = wpe / math.sqrt(mass_ratio)
    wgi = wge / mass_ratio
    r = [mass_ratio, 1.0]
    n0 = n_ppc
    q = [+math.sqrt(r[0] / (4 * math.pi * n0 / delx ** 2)) * wpi,
         -math.sqrt(r[1] / (4 * math.pi * n0 / delx ** 2)) * wpe]
    b0 = r[0] * c / q[0] * wgi
    return dict(nx=nx, ny=ny, nranks=nranks, n0=n0, np=int(math.ceil(n_ppc * nx * cap_factor)),
                nsp=2, delx=delx, delt=delt, c=c, gfac=gfac, q=q, r=r, b0=b0,
                vti=v_thi, vte=v_the, t_ani=t_ani, nxgs=2, nygs=2)


class World:
    """N y-slabs of the reference's data model in one address space."""

    def __init__(self, prm, fast=False):
        self.prm = dict(prm)
        self.lib = load(fast)
        cfg = OrcConfig()
        cfg.nx, cfg.ny, cfg.nxgs, cfg.nygs = prm["nx"], prm["ny"], prm.get("nxgs", 2), prm.get("nygs", 2)
        cfg.nranks, cfg.np, cfg.nsp, cfg.bc = prm["nranks"], prm["np"], prm["nsp"], prm.get("bc", 0)
        cfg.delx, cfg.delt, cfg.c, cfg.gfac = prm["delx"], prm["delt"], prm["c"], prm["gfac"]
        for s in range(prm["nsp"]):
            cfg.q[s] = prm["q"][s]
            cfg.r[s] = prm["r"][s]
        self.cfg = cfg
        self.h = self.lib.orc_create(C.byref(cfg))
        if not self.h:
            raise RuntimeError("orc_create failed")
        self.nranks = prm["nranks"]
        self.nx, self.ny, self.np_cap, self.nsp = prm["nx"], prm["ny"], prm["np"], prm["nsp"]
        self.nxgs, self.nygs = cfg.nxgs, cfg.nygs
        self.nxge, self.nyge = cfg.nxgs + cfg.nx - 1, cfg.nygs + cfg.ny - 1

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def bounds(self, rank):
        a, b = C.c_int32(), C.c_int32()
        self.lib.orc_bounds(self.h, rank, C.byref(a), C.byref(b))
        return a.value, b.value

    def array(self, rank, which):
        """Zero-copy numpy view, shaped in C order = reversed Fortran order."""
        n = C.c_int64()
        ptr = self.lib.orc_array(self.h, rank, which, C.byref(n))
        nys, nye = self.bounds(rank)
        nyl = nye - nys + 1
        if which in (NP2, CUMCNT):
            a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_int32)), shape=(n.value,))
            return a.reshape((self.nsp, nyl) if which == NP2 else (self.nsp, nyl, self.nx + 1))
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n.value,))
        shp = {UP: (self.nsp, nyl, self.np_cap, 6), GP: (self.nsp, nyl, self.np_cap, 6),
               UF: (nyl + 4, self.nx + 4, 6), DF: (nyl + 4, self.nx + 4, 6),
               UJ: (nyl + 4, self.nx + 4, 3), GKL: (nyl, self.nx, 3),
               MOM: (self.nsp, nyl + 2, self.nx + 2, 7)}[which]
        return a.reshape(shp)

    # stages ---------------------------------------------------------------
    def ic_weibel(self, seed=20260117):
        p = self.prm
        self.lib.orc_ic_weibel(self.h, seed, p["n0"], p["vti"], p["vte"], p["t_ani"], p["b0"])

    def particle_solv(self): self.lib.orc_particle_solv(self.h)
    def ele_cur(self): self.lib.orc_ele_cur(self.h)
    def bc_curre(self): self.lib.orc_bc_curre(self.h)
    def field_fdtd_i(self): return self.lib.orc_field_fdtd_i(self.h)
    def bc_particle_x(self): self.lib.orc_bc_particle_x(self.h)
    def bc_particle_y(self): return self.lib.orc_bc_particle_y(self.h)
    def bc_injection(self, u0): self.lib.orc_bc_injection(self.h, u0)
    def set_u_inject(self, u0): self.lib.orc_set_u_inject(self.h, u0)
    def sort_bucket(self): self.lib.orc_sort_bucket(self.h)
    def mom_accl(self): self.lib.orc_mom_accl(self.h)
    def mom_nvt(self): self.lib.orc_mom_nvt(self.h)
    def bc_mom(self): self.lib.orc_bc_mom(self.h)

    def step(self, n=1):
        e = self.lib.orc_step(self.h, n)
        if e:
            raise RuntimeError("oracle step failed: code %d" % e)

    def stage_times(self, reset=True):
        out = (C.c_double * 5)()
        self.lib.orc_stage_times(self.h, out, int(reset))
        return list(out)

    def cg_iters(self):
        out = (C.c_int32 * 3)()
        self.lib.orc_cg_iters(self.h, out)
        return list(out)

    def energy(self):
        out = (C.c_double * (self.nsp + 2))()
        self.lib.orc_energy(self.h, out)
        return np.array(list(out))

    def gauss_residual(self):
        sc = C.c_double()
        r = self.lib.orc_gauss_residual(self.h, C.byref(sc))
        return r, sc.value

    # helpers ----------------------------------------------------------------
    def particles_by_id(self, which=UP):
        """All active particles of all ranks as (ids, isp, rec[n,5]) sorted by (isp, id)."""
        ids, sp, rec = [], [], []
        for rk in range(self.nranks):
            a = self.array(rk, which)
            n2 = self.array(rk, NP2)
            for isp in range(self.nsp):
                for jl in range(a.shape[1]):
                    n = n2[isp, jl]
                    blk = a[isp, jl, :n, :]
                    ids.append(blk[:, 5].copy().view(np.int64))
                    sp.append(np.full(n, isp, np.int64))
                    rec.append(blk[:, :5].copy())
        ids = np.concatenate(ids)
        sp = np.concatenate(sp)
        rec = np.concatenate(rec)
        order = np.lexsort((ids, sp))
        return ids[order], sp[order], rec[order]

    def global_field(self, which=UF):
        """Interior of a grid array assembled over ranks: (ny, nx, ncomp)."""
        rows = []
        for rk in range(self.nranks):
            a = self.array(rk, which)
            rows.append(a[2:-2, 2:-2, :] if which != GKL else a)
        return np.concatenate(rows, axis=0)

    def cell_counts(self):
        """Per-cell particle counts from cumcnt: (nsp, ny, nx)."""
        out = []
        for rk in range(self.nranks):
            cc = self.array(rk, CUMCNT)
            out.append(np.diff(cc, axis=2))
        return np.concatenate(out, axis=1)
