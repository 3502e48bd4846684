"""The ring path on ONE GPU: N contexts of one process (one y slab each), driven by N host threads, exchanging through the
in-process loopback transport instead of NCCL (wm_loopback_create / wm_comm_init_loopback).  The same call sites as the NCCL
path -- leavers packed by the fused kernel and appended by k_incoming_append, the J fold and refresh of bc__curre, the
delta-field halos of bc__dfield, the CG halos and all-reduces (host-loop CG), the moments fold, the Gauss residual over the
ring -- against the oracle's N-slab world (common/boundary_periodic.f90:99-248,357-636, common/mpi_set.f90:36-47).
This is what makes multi-rank parity visible on a one-GPU box; tests/test_gpu_multi.py is the same on real GPUs.  -m gpu."""
import threading

import numpy as np
import pytest

import oracle_lib as O
from helpers import flatten_by_id, make_shock_world, make_wall_world, particle_err, rel_to_max

pytestmark = pytest.mark.gpu


def _run_ring(prm, w, nsteps, kind, env_exact=False):
    import wumingpic2d_b200 as wm
    n = prm["nranks"]
    init = [dict(up=w.array(r, O.UP).copy(), np2=w.array(r, O.NP2).copy(), cum=w.array(r, O.CUMCNT).copy(), uf=w.array(r, O.UF).copy())
            for r in range(n)]
    # the oracle's states after every step, per rank (its arrays are views that the next step overwrites)
    ref = []
    for it in range(nsteps):
        w.step(1)
        ref.append([dict(up=w.array(r, O.UP).copy(), np2=w.array(r, O.NP2).copy(), cum=w.array(r, O.CUMCNT).copy(),
                         uf=w.array(r, O.UF).copy(), cg=w.cg_iters()) for r in range(n)])
    w.mom_accl(); w.mom_nvt(); w.bc_mom()
    mom_ref = [w.array(r, O.MOM).copy() for r in range(n)]
    e_ref = w.energy()
    group = wm.LoopbackGroup(n)
    errs, energies, gauss = [None] * n, [None] * n, [None] * n

    def rank_main(r):
        try:
            nys, nye = w.bounds(r)
            c = wm.Context.from_params(prm, nys=nys, nye=nye, nrank=r, nsize=n, device=0,
                                       flags=wm.WM_FLAG_EXACT_PUSH if env_exact else 0)
            c.comm_init_loopback(group)
            if kind == "shock":
                c.set_u_inject(prm["u0"])
            c.upload_particles_sorted(init[r]["up"], init[r]["np2"], init[r]["cum"])
            c.upload_field(init[r]["uf"])
            for it in range(nsteps):
                c.step(1)
                assert c.cg_iters() == ref[it][r]["cg"], (it, c.cg_iters(), ref[it][r]["cg"])
                up, np2, cum = c.download_particles()
                assert np.array_equal(cum, ref[it][r]["cum"]), "per-cell counts differ on rank %d step %d" % (r, it)
                a, b = flatten_by_id(up, np2), flatten_by_id(ref[it][r]["up"], ref[it][r]["np2"])
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
                ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
                tol = 1e-12 if it == 0 else 1e-10
                assert ex <= tol and eu <= tol, (r, it, ex, eu)
                assert rel_to_max(c.download_field(), ref[it][r]["uf"]).max() <= tol
            mom = c.moments()
            assert rel_to_max(mom[:, 1:-1, 1:-1], mom_ref[r][:, 1:-1, 1:-1]).max() <= 1e-10
            energies[r] = c.energy()
            if kind == "periodic":
                gauss[r] = c.gauss_residual()
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
    assert np.allclose(np.sum(energies, axis=0), e_ref, rtol=1e-10, atol=0)
    if kind == "periodic":
        res = max(g[0] for g in gauss)
        scale = max(g[1] for g in gauss)
        assert res <= 1e-11 * scale, (res, scale)


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_loopback_ring_periodic(nranks):
    """ny not divisible by the rank count: the remainder rule of mpi_set.f90:37-41; four steps + moments + Gauss law"""
    prm = O.weibel_params(40, 8 * nranks + 3, 12, nranks=nranks)
    w = O.World(prm)
    w.ic_weibel(20260117)
    _run_ring(prm, w, 4, "periodic")
    w.close()


@pytest.mark.parametrize("kind", ["reconnection", "shock"])
def test_loopback_ring_walls(kind):
    if kind == "reconnection":
        prm, w = make_wall_world(40, 19, 8, nranks=2)
    else:
        prm, w = make_shock_world(40, 19, 8, u0=-0.3, nranks=2)
    _run_ring(prm, w, 4, kind)
    w.close()


def test_loopback_ring_exact_path():
    """the tag + scatter sort and the stage-call exchange (migrate with counts first, scatter of the arrivals)"""
    prm = O.weibel_params(40, 19, 10, nranks=2)
    w = O.World(prm)
    w.ic_weibel(20260117)
    _run_ring(prm, w, 3, "periodic", env_exact=True)
    w.close()
