"""CPU tests (-m "not gpu"): the oracle against invariants, and that the C-ABI library loads and
exports every symbol include/wumingpic2d.h declares (no compute calls without a GPU)."""
import os
import re

import numpy as np
import pytest

import oracle_lib as O
from helpers import flatten_by_id, make_world

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(wmlib):
    hdr = open(os.path.join(ROOT, "include", "wumingpic2d.h")).read()
    names = set(re.findall(r"\b(wm_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 30
    for n in sorted(names):
        assert hasattr(wmlib, n), "libwumingpic2d.so does not export " + n
    from wumingpic2d_b200.api import EXPORTS
    assert names == set(EXPORTS)


def test_no_cpu_fallback(wmlib):
    """Without a device the product path must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import wumingpic2d_b200 as wm
    prm = O.weibel_params(16, 16, 2)
    with pytest.raises(wm.WmError, match="no CUDA device"):
        wm.Context.from_params(prm)


def test_product_does_not_reference_oracle():
    for d, _, files in os.walk(os.path.join(ROOT, "wumingpic2d_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp")):
                src = open(os.path.join(d, f)).read()
                for needle in ("oracle_lib", "wm_oracle", "orc_", "oracle/", "import oracle", "from oracle"):
                    assert needle not in src, "%s references the test oracle (%s)" % (f, needle)


def test_oracle_gauss_law_and_energy():
    """Discrete Gauss law holds to roundoff through deposit + field solve + boundaries + sort;
    total energy stays within 1e-3 over 40 steps (implicit scheme, gfac=0.501)."""
    prm, w = make_world(32, 32, 10)
    e0 = w.energy().sum()
    r0, s0 = w.gauss_residual()
    assert r0 <= 1e-13 * s0
    for _ in range(4):
        w.step(10)
        r, s = w.gauss_residual()
        assert r <= 1e-12 * s
        assert all(1 <= k < 100 for k in w.cg_iters())
    assert abs(w.energy().sum() - e0) <= 1e-3 * e0
    cnt = w.cell_counts()
    assert cnt.sum() == 2 * 32 * 32 * 10
    w.close()


def test_oracle_sort_invariants():
    prm, w = make_world(24, 12, 6, steps=5)
    up, np2, cum = w.array(0, O.UP), w.array(0, O.NP2), w.array(0, O.CUMCNT)
    for isp in range(2):
        for jl in range(12):
            n = np2[isp, jl]
            xs = up[isp, jl, :n, 0].astype(np.int64) - 2
            ys = up[isp, jl, :n, 1].astype(np.int64) - 2
            assert np.all(ys == jl)
            assert np.all(np.diff(xs) >= 0)
            assert np.array_equal(np.bincount(xs, minlength=24), np.diff(cum[isp, jl]))
            assert cum[isp, jl, 0] == 0 and cum[isp, jl, -1] == n
    w.close()


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_oracle_nrank_equals_one_rank(nranks):
    """The y-slab decomposition (common/mpi_set.f90:36-47) does not change the answer."""
    _, w1 = make_world(32, 30, 8, steps=12)
    _, wn = make_world(32, 30, 8, nranks=nranks, steps=12)
    i1, s1, r1 = w1.particles_by_id()
    i2, s2, r2 = wn.particles_by_id()
    assert np.array_equal(i1, i2) and np.array_equal(s1, s2)
    assert np.abs(r1 - r2).max() <= 1e-12
    assert np.abs(w1.global_field() - wn.global_field()).max() <= 1e-13
    assert np.array_equal(w1.cell_counts(), wn.cell_counts())
    assert w1.cg_iters() == wn.cg_iters()
    w1.close(); wn.close()


def test_oracle_rng_is_decomposition_independent():
    _, w1 = make_world(16, 16, 4)
    _, w4 = make_world(16, 16, 4, nranks=4)
    a, b = w1.particles_by_id(), w4.particles_by_id()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])
    u = a[2][:, 2:]
    assert abs(u[:, 0].std() - 0.1) < 5e-3 and abs(u[:, 2].std() - 0.5) < 2.5e-2
    w1.close(); w4.close()


@pytest.mark.parametrize("nranks", [1, 3])
def test_oracle_reconnection_walls(nranks):
    """proj/reconnection boundary module in the oracle: particles stay between the reflecting walls,
    their number is conserved, total energy is conserved to 5e-3, and the slab decomposition does
    not change the answer."""
    from helpers import make_wall_world
    prm, w = make_wall_world(32, 18, 8, nranks=nranks)
    _, w1 = make_wall_world(32, 18, 8, nranks=1)
    e0 = w.energy().sum()
    w.step(15)
    w1.step(15)
    ids, sp, rec = w.particles_by_id()
    assert len(ids) == 2 * 8 * (32 - 3) * 18
    assert rec[:, 0].min() >= prm["nxgs"] + 1 and rec[:, 0].max() < prm["nxgs"] + 32 - 2
    assert abs(w.energy().sum() - e0) <= 5e-3 * e0   # not an equilibrium: fields build up from noise
    i1, s1, r1 = w1.particles_by_id()
    assert np.array_equal(ids, i1) and np.abs(rec - r1).max() <= 1e-12
    assert np.abs(w.global_field() - w1.global_field()).max() <= 1e-13
    assert w.cg_iters() == w1.cg_iters()
    w.close(); w1.close()


@pytest.mark.parametrize("nranks", [1, 2])
def test_oracle_shock_injection(nranks):
    """proj/shock boundary module in the oracle.  Known answers for bc__injection (boundary_shock.f90:255-297): a
    particle left of nxs+1 is mirrored there with u -> -u; a particle beyond xend = nxe*delx + v0*delt is mirrored
    at xend with ux -> 2 u0 - ux, uy, uz -> -uy, -uz; everything else is untouched.  Then a run: particles stay
    between the walls, their number is conserved, the discrete Gauss law residual stays at roundoff although
    particles are reflected (the injection precedes the deposit, proj/shock/app.f90:112-113), and the slab
    decomposition does not change the answer."""
    from helpers import make_shock_world
    u0 = -0.3
    prm, w = make_shock_world(32, 16, 6, u0=u0, nranks=nranks)
    nxgs, nx = prm["nxgs"], prm["nx"]
    nxs, nxe = nxgs, nxgs + nx - 1
    v0 = u0 / np.sqrt(1 + u0 * u0)
    xend = nxe * 1.0 + v0 * prm["delt"]
    # ---- known answers on three hand-placed records of rank 0, row 0, species 0
    gp = w.array(0, O.GP)
    keep = gp[0, 0, :3].copy()
    gp[0, 0, 0, :5] = (nxs + 0.75, gp[0, 0, 0, 1], -0.4, 0.1, -0.2)       # left of the wall at nxs+1
    gp[0, 0, 1, :5] = (xend + 0.125, gp[0, 0, 1, 1], 0.05, 0.1, -0.2)     # beyond the injection wall
    gp[0, 0, 2, :5] = (xend - 0.125, gp[0, 0, 2, 1], 0.05, 0.1, -0.2)     # inside
    w.bc_injection(u0)
    assert np.array_equal(gp[0, 0, 0, :5], (2. * (nxs + 1) - (nxs + 0.75), keep[0, 1], 0.4, -0.1, 0.2))
    assert np.array_equal(gp[0, 0, 1, :5], (2. * xend - (xend + 0.125), keep[1, 1], 2. * u0 - 0.05, -0.1, 0.2))
    assert np.array_equal(gp[0, 0, 2, :5], (xend - 0.125, keep[2, 1], 0.05, 0.1, -0.2))
    gp[0, 0, :3] = keep
    # ---- a run
    _, w1 = make_shock_world(32, 16, 6, u0=u0, nranks=1)
    r0, s0 = w.gauss_residual()
    w.step(12)
    w1.step(12)
    ids, sp, rec = w.particles_by_id()
    assert len(ids) == 2 * 6 * (32 - 2) * 16
    assert rec[:, 0].min() >= nxs + 1 and rec[:, 0].max() <= xend
    r, s = w.gauss_residual()
    assert abs(r - r0) <= 1e-11 * max(s, 1.0)
    i1, s1, r1 = w1.particles_by_id()
    assert np.array_equal(ids, i1) and np.abs(rec - r1).max() <= 1e-12
    assert np.abs(w.global_field() - w1.global_field()).max() <= 1e-13
    w.close(); w1.close()


def test_library_is_built_from_this_tree():
    """The .so is git-ignored and travels with the tree: load_library() refuses a binary whose compiled-in source hash
    (wm_source_hash) differs from the hash of csrc/ + include/ as they are now."""
    import wumingpic2d_b200 as wm
    from wumingpic2d_b200.build import source_hash
    lib = wm.load_library()
    assert lib.wm_source_hash().decode() == source_hash()
    assert lib.wm_version() >= 200
