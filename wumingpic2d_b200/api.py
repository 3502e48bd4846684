"""ctypes binding of include/wumingpic2d.h.

Method names follow the reference's Fortran module procedures (module__procedure), so a
driver loop reads like proj/weibel/app.f90:100-107:

    ctx.particle__solv(); ctx.field__fdtd_i(); ctx.bc__particle_x(); ctx.bc__particle_y();
    ctx.sort__bucket()

Host arrays are numpy float64/int32 buffers in the reference's Fortran layout (see the header).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
WM_NSP_MAX = 2
WM_BC_PERIODIC = 0
WM_BC_RECONNECTION = 1
WM_BC_SHOCK = 2
WM_FLAG_EXACT_PUSH = 1


class WmError(RuntimeError):
    pass


class WmConfig(C.Structure):
    _fields_ = [
        ("ndim", C.c_int32), ("np", C.c_int32), ("nsp", C.c_int32),
        ("nxgs", C.c_int32), ("nxge", C.c_int32), ("nygs", C.c_int32), ("nyge", C.c_int32),
        ("nys", C.c_int32), ("nye", C.c_int32), ("nrank", C.c_int32), ("nsize", C.c_int32),
        ("bc", C.c_int32), ("device", C.c_int32), ("flags", C.c_int32),
        ("delx", C.c_double), ("delt", C.c_double), ("c", C.c_double), ("gfac", C.c_double),
        ("q", C.c_double * WM_NSP_MAX), ("r", C.c_double * WM_NSP_MAX),
        ("capacity", C.c_int64),
    ]


def library_path():
    """the in-tree library; WM_LIB selects another build of the same sources (compiler-flag experiments, scripts/build_variants.sh)"""
    return os.environ.get("WM_LIB") or os.path.join(HERE, "libwumingpic2d.so")


_lib = None

EXPORTS = [
    "wm_last_error", "wm_version", "wm_source_hash", "wm_create", "wm_destroy", "wm_comm_unique_id", "wm_comm_init",
    "wm_upload_particles", "wm_upload_particles_sorted", "wm_upload_field", "wm_download_particles",
    "wm_download_gp", "wm_download_field", "wm_download_current", "wm_download_dfield",
    "wm_particle_counts", "wm_particle__solv", "wm_field__ele_cur", "wm_boundary__curre",
    "wm_field__fdtd_i", "wm_boundary__particle_x", "wm_boundary__particle_y", "wm_boundary__injection",
    "wm_set_u_inject", "wm_set_xrange", "wm_append_particles", "wm_sort__bucket",
    "wm_step", "wm_host_step", "wm_host_steps", "wm_host_pipe_chunks", "wm_host_pipe_plan", "wm_host_register", "wm_host_unregister", "wm_loopback_create", "wm_loopback_destroy", "wm_comm_init_loopback", "wm_cg_path", "wm_cg_plan", "wm_fp64_peak", "wm_ic_harris", "wm_ic_shock", "wm_shock_inject", "wm_shock_relocate", "wm_xrange", "wm_host_particle__solv", "wm_host_sort__bucket", "wm_cg_iters",
    "wm_energy", "wm_gauss_residual", "wm_moments", "wm_mom_calc__accl", "wm_mom_calc__nvt", "wm_boundary__mom", "wm_ic_weibel", "wm_timing", "wm_synchronize", "wm_layout_rebuilds",
]


def _preload_nccl():
    """libwumingpic2d.so needs libnccl.so.2.  PyTorch bundles a newer NCCL under the same soname; if
    the system copy were loaded first, a later `import torch` in the same process would bind to it
    and miss symbols.  So load the torch-bundled copy first when there is one (the ABI we use --
    comm init, send/recv, all-reduce, groups -- is stable across 2.x)."""
    import sys
    for d in sys.path:
        cand = os.path.join(d, "nvidia", "nccl", "lib", "libnccl.so.2")
        if os.path.exists(cand):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return cand
            except OSError:
                pass
    return None


def load_library():
    """Load libwumingpic2d.so; raises if it has not been built (python -m wumingpic2d_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise WmError("%s not found: build it with `python -m wumingpic2d_b200.build` "
                      "(there is no CPU fallback)" % path)
    _preload_nccl()
    lib = C.CDLL(path)
    P, D, I32 = C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int32)
    lib.wm_last_error.restype = C.c_char_p
    lib.wm_source_hash.restype = C.c_char_p
    if not os.environ.get("WM_SKIP_HASH_CHECK"):
        from .build import source_hash
        built, tree = lib.wm_source_hash().decode(), source_hash()
        if built != tree:
            raise WmError("%s was built from other sources (binary %s, tree %s): rebuild with `python -m wumingpic2d_b200.build`"
                          % (path, built, tree))
    lib.wm_create.argtypes = [C.POINTER(WmConfig), C.POINTER(P)]
    lib.wm_destroy.argtypes = [P]
    lib.wm_comm_unique_id.argtypes = [C.c_void_p]
    lib.wm_comm_init.argtypes = [P, C.c_void_p]
    lib.wm_loopback_create.argtypes = [C.c_int32, C.POINTER(P)]
    lib.wm_loopback_destroy.argtypes = [P]
    lib.wm_comm_init_loopback.argtypes = [P, P]
    lib.wm_upload_particles.argtypes = [P, D, I32]
    lib.wm_upload_particles_sorted.argtypes = [P, D, I32, I32]
    lib.wm_upload_field.argtypes = [P, D]
    lib.wm_download_particles.argtypes = [P, D, I32, I32]
    lib.wm_download_gp.argtypes = [P, D]
    lib.wm_download_field.argtypes = [P, D]
    lib.wm_download_current.argtypes = [P, D]
    lib.wm_download_dfield.argtypes = [P, D]
    lib.wm_particle_counts.argtypes = [P, C.POINTER(C.c_int64)]
    for n in ("wm_particle__solv", "wm_field__ele_cur", "wm_boundary__curre", "wm_field__fdtd_i",
              "wm_boundary__particle_x", "wm_boundary__particle_y", "wm_sort__bucket", "wm_synchronize"):
        getattr(lib, n).argtypes = [P]
    lib.wm_step.argtypes = [P, C.c_int32]
    lib.wm_boundary__injection.argtypes = [P, C.c_double]
    lib.wm_set_u_inject.argtypes = [P, C.c_double]
    lib.wm_set_xrange.argtypes = [P, C.c_int32, C.c_int32]
    lib.wm_append_particles.argtypes = [P, C.c_int32, C.c_int64, D]
    lib.wm_host_step.argtypes = [P, D, D, I32, I32]
    lib.wm_host_steps.argtypes = [P, D, D, I32, I32, C.c_int32]
    lib.wm_host_pipe_chunks.argtypes = [P]
    lib.wm_host_pipe_plan.argtypes = [C.c_int32, C.c_int32, I32, C.c_int32]
    lib.wm_host_register.argtypes = [C.c_void_p, C.c_size_t]
    lib.wm_host_unregister.argtypes = [C.c_void_p]
    lib.wm_host_particle__solv.argtypes = [P, D, D, D, I32, I32]
    lib.wm_host_sort__bucket.argtypes = [P, D, D, I32, I32]
    lib.wm_cg_iters.argtypes = [P, I32]
    lib.wm_cg_path.argtypes = [P, I32]
    lib.wm_fp64_peak.argtypes = [P, D]
    lib.wm_cg_plan.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int64, I32]
    lib.wm_energy.argtypes = [P, D]
    lib.wm_gauss_residual.argtypes = [P, C.POINTER(C.c_double)]
    lib.wm_moments.argtypes = [P, D]
    lib.wm_mom_calc__accl.argtypes = [P]
    lib.wm_mom_calc__nvt.argtypes = [P, D]
    lib.wm_boundary__mom.argtypes = [P, D]
    lib.wm_ic_weibel.argtypes = [P, C.c_uint64, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double]
    lib.wm_ic_harris.argtypes = [P, C.c_uint64, C.c_int32, C.c_int32] + [C.c_double] * 6
    lib.wm_ic_shock.argtypes = [P, C.c_uint64, C.c_int32, C.c_int32] + [C.c_double] * 7
    lib.wm_shock_inject.argtypes = [P, C.c_uint64, C.c_int32]
    lib.wm_shock_relocate.argtypes = [P, C.c_uint64, C.c_int32]
    lib.wm_xrange.argtypes = [P, I32]
    lib.wm_timing.argtypes = [P, D, C.POINTER(C.c_int64), C.c_int32]
    lib.wm_layout_rebuilds.argtypes = [P, C.POINTER(C.c_int64)]
    _lib = lib
    return lib


def cg_plan(nx, nyl, nsm=148, smem_max=227 * 1024 - 2048):
    """Block decomposition of the persistent CG kernel for an nx x nyl slab: (blocks in x, blocks in y, smem bytes, cells
    per thread)."""
    out = (C.c_int32 * 4)()
    lib = load_library()
    if lib.wm_cg_plan(nx, nyl, nsm, smem_max, out):
        raise WmError(lib.wm_last_error().decode())
    return tuple(out)


def host_pipe_plan(nyl, rows=16):
    """the schedule of a pipelined host step: list of (kind, a, b) -- 'push' / 'place' tile rows [a,b), 'down' rows [a,b), 'ring'"""
    lib = load_library()
    buf = np.zeros(3 * (4 * (nyl // 8 + 4) + 8), dtype=np.int32)
    n = lib.wm_host_pipe_plan(nyl, rows, buf.ctypes.data_as(C.POINTER(C.c_int32)), buf.size // 3)
    if n < 0:
        raise WmError("wm_host_pipe_plan(%d, %d) = %d" % (nyl, rows, n))
    names = ("push", "place", "down", "ring")
    return [(names[buf[3 * i]], int(buf[3 * i + 1]), int(buf[3 * i + 2])) for i in range(n)]


def host_register(a):
    """page-lock a numpy array for the host-array calls (wm_host_register); returns the array"""
    lib = load_library()
    if lib.wm_host_register(a.ctypes.data_as(C.c_void_p), a.nbytes) != 0:
        raise WmError(lib.wm_last_error().decode())
    return a


def host_unregister(a):
    lib = load_library()
    if lib.wm_host_unregister(a.ctypes.data_as(C.c_void_p)) != 0:
        raise WmError(lib.wm_last_error().decode())


def _d(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_int32))


class LoopbackGroup:
    """The in-process transport for several ranks on one device (wm_loopback_create): every rank's Context is attached with
    Context.comm_init_loopback(group) and driven by its own host thread."""

    def __init__(self, nsize):
        self.lib = load_library()
        self.h = C.c_void_p()
        if self.lib.wm_loopback_create(nsize, C.byref(self.h)):
            raise WmError(self.lib.wm_last_error().decode())
        self.nsize = nsize

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.lib.wm_loopback_destroy(self.h)
            self.h = None


class Context:
    """One GPU = one y-slab.  Constructor arguments are those of the reference's *__init calls
    (proj/weibel/app.f90:331-347) plus the ring position of common/mpi_set.f90."""

    def __init__(self, *, np_cap, nxgs, nxge, nygs, nyge, nys, nye, delx, delt, c, q, r, gfac,
                 ndim=6, nsp=2, nrank=0, nsize=1, bc=WM_BC_PERIODIC, device=-1, flags=0, capacity=0):
        self.lib = load_library()
        g = WmConfig()
        g.ndim, g.np, g.nsp = ndim, np_cap, nsp
        g.nxgs, g.nxge, g.nygs, g.nyge, g.nys, g.nye = nxgs, nxge, nygs, nyge, nys, nye
        g.nrank, g.nsize, g.bc, g.device, g.flags = nrank, nsize, bc, device, flags
        g.delx, g.delt, g.c, g.gfac = delx, delt, c, gfac
        for s in range(nsp):
            g.q[s] = q[s]
            g.r[s] = r[s]
        g.capacity = capacity
        self.cfg = g
        self.h = C.c_void_p()
        self._ck(self.lib.wm_create(C.byref(g), C.byref(self.h)))
        self.nx, self.nyl, self.nsp, self.np_cap = nxge - nxgs + 1, nye - nys + 1, nsp, np_cap

    @classmethod
    def from_params(cls, prm, nys=None, nye=None, nrank=0, **kw):
        """prm: dict with nx, ny, np, nsp, delx, delt, c, gfac, q, r (and optionally nxgs, nygs)."""
        nxgs, nygs = prm.get("nxgs", 2), prm.get("nygs", 2)
        nyge = nygs + prm["ny"] - 1
        return cls(np_cap=prm["np"], nxgs=nxgs, nxge=nxgs + prm["nx"] - 1, nygs=nygs, nyge=nyge,
                   nys=nygs if nys is None else nys, nye=nyge if nye is None else nye,
                   delx=prm["delx"], delt=prm["delt"], c=prm["c"], q=prm["q"], r=prm["r"],
                   gfac=prm["gfac"], nsp=prm["nsp"], nrank=nrank, bc=kw.pop("bc", prm.get("bc", WM_BC_PERIODIC)), **kw)

    def _ck(self, rc):
        if rc:
            raise WmError(self.lib.wm_last_error().decode())

    def close(self):
        if getattr(self, "h", None) and self.h:
            self.lib.wm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- shapes of the host arrays (C order = reversed Fortran order)
    def shape_up(self): return (self.nsp, self.nyl, self.np_cap, 6)
    def shape_uf(self): return (self.nyl + 4, self.nx + 4, 6)
    def shape_uj(self): return (self.nyl + 4, self.nx + 4, 3)
    def shape_np2(self): return (self.nsp, self.nyl)
    def shape_cumcnt(self): return (self.nsp, self.nyl, self.nx + 1)
    def shape_mom(self): return (self.nsp, self.nyl + 2, self.nx + 2, 7)

    # ---- communicator
    def comm_unique_id(self):
        buf = C.create_string_buffer(128)
        self._ck(self.lib.wm_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, id128):
        self._ck(self.lib.wm_comm_init(self.h, C.c_char_p(id128)))

    # ---- residency
    def upload_particles(self, up, np2): self._ck(self.lib.wm_upload_particles(self.h, _d(up), _i(np2)))
    def upload_particles_sorted(self, up, np2, cumcnt):
        self._ck(self.lib.wm_upload_particles_sorted(self.h, _d(up), _i(np2), _i(cumcnt)))
    def upload_field(self, uf): self._ck(self.lib.wm_upload_field(self.h, _d(uf)))

    def download_particles(self, up=None, want_up=True):
        if up is None and want_up:
            up = np.zeros(self.shape_up())
        np2 = np.zeros(self.shape_np2(), np.int32)
        cum = np.zeros(self.shape_cumcnt(), np.int32)
        self._ck(self.lib.wm_download_particles(self.h, _d(up), _i(np2), _i(cum)))
        return up, np2, cum

    def download_gp(self):
        gp = np.zeros(self.shape_up())
        self._ck(self.lib.wm_download_gp(self.h, _d(gp)))
        return gp

    def download_field(self):
        uf = np.zeros(self.shape_uf())
        self._ck(self.lib.wm_download_field(self.h, _d(uf)))
        return uf

    def download_current(self):
        uj = np.zeros(self.shape_uj())
        self._ck(self.lib.wm_download_current(self.h, _d(uj)))
        return uj

    def download_dfield(self):
        df = np.zeros(self.shape_uf())
        self._ck(self.lib.wm_download_dfield(self.h, _d(df)))
        return df

    def particle_counts(self):
        n = (C.c_int64 * WM_NSP_MAX)()
        self._ck(self.lib.wm_particle_counts(self.h, n))
        return [n[s] for s in range(self.nsp)]

    # ---- the reference's module procedures, device resident
    def particle__solv(self): self._ck(self.lib.wm_particle__solv(self.h))
    def field__ele_cur(self): self._ck(self.lib.wm_field__ele_cur(self.h))
    def bc__curre(self): self._ck(self.lib.wm_boundary__curre(self.h))
    def field__fdtd_i(self): self._ck(self.lib.wm_field__fdtd_i(self.h))
    def bc__particle_x(self): self._ck(self.lib.wm_boundary__particle_x(self.h))
    def bc__particle_y(self): self._ck(self.lib.wm_boundary__particle_y(self.h))
    def bc__injection(self, u0): self._ck(self.lib.wm_boundary__injection(self.h, u0))
    def set_u_inject(self, u0): self._ck(self.lib.wm_set_u_inject(self.h, u0))
    def set_xrange(self, nxs, nxe): self._ck(self.lib.wm_set_xrange(self.h, nxs, nxe))

    def append_particles(self, isp, rec):
        rec = np.ascontiguousarray(rec, dtype=np.float64).reshape(-1, 6)
        self._ck(self.lib.wm_append_particles(self.h, isp, rec.shape[0], _d(rec)))
    def sort__bucket(self): self._ck(self.lib.wm_sort__bucket(self.h))
    def step(self, n=1): self._ck(self.lib.wm_step(self.h, n))

    # ---- host-array (drop-in) calls
    def host_step(self, up, uf, np2, cumcnt):
        self._ck(self.lib.wm_host_step(self.h, _d(up), _d(uf), _i(np2), _i(cumcnt)))

    def host_pipe_chunks(self):
        """chunks of the last host_step (0 = upload, step, download one after the other)"""
        return int(self.lib.wm_host_pipe_chunks(self.h))

    def host_steps(self, up, uf, np2, cumcnt, nsteps):
        self._ck(self.lib.wm_host_steps(self.h, _d(up), _d(uf), _i(np2), _i(cumcnt), nsteps))

    def host_particle__solv(self, gp, up, uf, cumcnt, np2):
        self._ck(self.lib.wm_host_particle__solv(self.h, _d(gp), _d(up), _d(uf), _i(cumcnt), _i(np2)))

    def host_sort__bucket(self, gp_out, up_in, cumcnt, np2):
        self._ck(self.lib.wm_host_sort__bucket(self.h, _d(gp_out), _d(up_in), _i(cumcnt), _i(np2)))

    # ---- diagnostics
    def cg_iters(self):
        out = (C.c_int32 * 3)()
        self._ck(self.lib.wm_cg_iters(self.h, out))
        return list(out)

    def fp64_peak(self):
        v = C.c_double()
        self._ck(self.lib.wm_fp64_peak(self.h, C.byref(v)))
        return v.value

    def comm_init_loopback(self, group):
        self._ck(self.lib.wm_comm_init_loopback(self.h, group.h))

    def cg_path(self):
        v = C.c_int32()
        self._ck(self.lib.wm_cg_path(self.h, C.byref(v)))
        return v.value

    def energy(self):
        out = (C.c_double * (self.nsp + 2))()
        self._ck(self.lib.wm_energy(self.h, out))
        return np.array(list(out))

    def gauss_residual(self):
        out = (C.c_double * 2)()
        self._ck(self.lib.wm_gauss_residual(self.h, out))
        return out[0], out[1]

    def moments(self):
        mom = np.zeros(self.shape_mom())
        self._ck(self.lib.wm_moments(self.h, _d(mom)))
        return mom

    def mom_calc__accl(self): self._ck(self.lib.wm_mom_calc__accl(self.h))

    def mom_calc__nvt(self):
        mom = np.zeros(self.shape_mom())
        self._ck(self.lib.wm_mom_calc__nvt(self.h, _d(mom)))
        return mom

    def bc__mom(self, mom): self._ck(self.lib.wm_boundary__mom(self.h, _d(mom)))

    def ic_weibel(self, seed, n0, vti, vte, t_ani, b0):
        self._ck(self.lib.wm_ic_weibel(self.h, seed, n0, vti, vte, t_ani, b0))

    def ic_harris(self, seed, nbg, ncs, lcs, vti, vte, b0, rtemp, e1=0.12):
        self._ck(self.lib.wm_ic_harris(self.h, seed, nbg, ncs, lcs, vti, vte, b0, rtemp, e1))

    def ic_shock(self, seed, n0, nxe, v0, vti, vte, b0, theta_bn, phi_bn, l_damp_ini):
        self._ck(self.lib.wm_ic_shock(self.h, seed, n0, nxe, v0, vti, vte, b0, theta_bn, phi_bn, l_damp_ini))

    def shock_inject(self, seed, it): self._ck(self.lib.wm_shock_inject(self.h, seed, it))
    def shock_relocate(self, seed, it): self._ck(self.lib.wm_shock_relocate(self.h, seed, it))

    def xrange(self):
        out = (C.c_int32 * 2)()
        self._ck(self.lib.wm_xrange(self.h, out))
        return out[0], out[1]

    def rebuilds(self):
        n = C.c_int64()
        self._ck(self.lib.wm_layout_rebuilds(self.h, C.byref(n)))
        return n.value

    def timing(self, reset=True):
        ms = (C.c_double * 5)()
        n = C.c_int64()
        self._ck(self.lib.wm_timing(self.h, ms, C.byref(n), int(reset)))
        return list(ms), n.value

    def synchronize(self): self._ck(self.lib.wm_synchronize(self.h))
