"""Multi-GPU parity worker: launched by tests/test_gpu_multi.py under torch.distributed.run, one
rank per GPU.  Every rank builds the same global oracle world split into WORLD_SIZE y-slabs
(common/mpi_set.f90:36-47), uploads its own slab, steps, and compares its slab with the oracle's:
per-cell counts bit-exact, particles by ID and fields <= tolerance, equal CG iteration counts,
moments after the ring fold.  Exit code 0 = parity holds on this rank."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import oracle_lib as O  # noqa: E402
from helpers import flatten_by_id, particle_err, rel_to_max  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    import wumingpic2d_b200 as wm
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # ny is not divisible by the rank count: exercises the remainder rule of mpi_set.f90:37-41
    prm = O.weibel_params(40, 8 * world + 3, 12, nranks=world)
    w = O.World(prm)
    w.ic_weibel(20260117)
    nys, nye = w.bounds(rank)
    ctx = wm.Context.from_params(prm, nys=nys, nye=nye, nrank=rank, nsize=world, device=local)
    ids = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0])
    ctx.upload_particles_sorted(w.array(rank, O.UP).copy(), w.array(rank, O.NP2).copy(), w.array(rank, O.CUMCNT).copy())
    ctx.upload_field(w.array(rank, O.UF).copy())
    ok = True
    for it in range(nsteps):
        w.step(1)
        ctx.step(1)
        up, np2, cum = ctx.download_particles()
        assert ctx.cg_iters() == w.cg_iters(), (ctx.cg_iters(), w.cg_iters())
        assert np.array_equal(cum, w.array(rank, O.CUMCNT)), "per-cell counts differ on rank %d step %d" % (rank, it)
        a, b = flatten_by_id(up, np2), flatten_by_id(w.array(rank, O.UP), w.array(rank, O.NP2))
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
        ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
        tol = 1e-12 if it == 0 else 1e-10
        assert ex <= tol and eu <= tol, (it, ex, eu)
        assert rel_to_max(ctx.download_field(), w.array(rank, O.UF)).max() <= tol
    # the stage calls one by one + moments with the ring fold (boundary_periodic.f90:571-636)
    w.step(1)
    ctx.particle__solv(); ctx.field__fdtd_i(); ctx.bc__particle_x(); ctx.bc__particle_y(); ctx.sort__bucket()
    up, np2, cum = ctx.download_particles()
    assert np.array_equal(cum, w.array(rank, O.CUMCNT))
    assert rel_to_max(ctx.download_field(), w.array(rank, O.UF)).max() <= 1e-10
    w.mom_accl(); w.mom_nvt(); w.bc_mom()
    mom = ctx.moments()
    ref = w.array(rank, O.MOM)
    assert rel_to_max(mom[:, 1:-1, 1:-1], ref[:, 1:-1, 1:-1]).max() <= 1e-10
    # energy: this rank's share sums to the oracle's global value
    e = torch.tensor(ctx.energy(), dtype=torch.float64, device="cuda")
    dist.all_reduce(e)
    assert np.allclose(e.cpu().numpy(), w.energy(), rtol=1e-10, atol=0)
    ctx.close()
    w.close()
    # ---- the pipelined host step (wm_host_step: rows in chunks, upload / pass / download side by side) on the ring: the leavers
    #      of the edge rows cross to the neighbour GPUs at the end of the pass, the CG exchanges inside k_cg_persist
    if os.environ.get("WM_INPLACE") != "0":
        os.environ["WM_HOSTPIPE_ROWS"] = "8"
        prm = O.weibel_params(40, 24 * world + 3, 10, nranks=world)
        w = O.World(prm)
        w.ic_weibel(20260117)
        nys, nye = w.bounds(rank)
        ctx = wm.Context.from_params(prm, nys=nys, nye=nye, nrank=rank, nsize=world, device=local)
        ids = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0])
        up, uf = w.array(rank, O.UP).copy(), w.array(rank, O.UF).copy()
        np2, cum = w.array(rank, O.NP2).copy(), w.array(rank, O.CUMCNT).copy()
        for it in range(4):
            w.step(1)
            ctx.host_step(up, uf, np2, cum)
            assert ctx.host_pipe_chunks() >= 3, ctx.host_pipe_chunks()
            assert ctx.cg_iters() == w.cg_iters(), (ctx.cg_iters(), w.cg_iters())
            assert np.array_equal(np2, w.array(rank, O.NP2)) and np.array_equal(cum, w.array(rank, O.CUMCNT)), \
                "host step: counts differ on rank %d step %d" % (rank, it)
            a, b = flatten_by_id(up, np2), flatten_by_id(w.array(rank, O.UP), w.array(rank, O.NP2))
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
            ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
            tol = 1e-12 if it == 0 else 1e-10
            assert ex <= tol and eu <= tol, ("host step", it, ex, eu)
            assert rel_to_max(uf, w.array(rank, O.UF)).max() <= tol
        ctx.close()
        w.close()
    # ---- the wall boundary modules on the ring: reconnection (reflecting / conducting walls) and shock (injection wall,
    #      bc__injection before the deposit), both periodic in y over the ranks
    from helpers import make_shock_world, make_wall_world
    for kind in ("reconnection", "shock"):
        if kind == "shock" and os.environ.get("WM_INPLACE") == "0":
            continue   # WM_BC_SHOCK runs on k_fused_sm (in-place sort) and on the exact path only
        if kind == "reconnection":
            prm, w = make_wall_world(40, 8 * world + 3, 8, nranks=world)
        else:
            prm, w = make_shock_world(40, 8 * world + 3, 8, u0=-0.3, nranks=world)
        nys, nye = w.bounds(rank)
        ctx = wm.Context.from_params(prm, nys=nys, nye=nye, nrank=rank, nsize=world, device=local)
        ids = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0])
        if kind == "shock":
            ctx.set_u_inject(-0.3)
        ctx.upload_particles_sorted(w.array(rank, O.UP).copy(), w.array(rank, O.NP2).copy(), w.array(rank, O.CUMCNT).copy())
        ctx.upload_field(w.array(rank, O.UF).copy())
        for it in range(nsteps):
            w.step(1)
            ctx.step(1)
            assert ctx.cg_iters() == w.cg_iters(), (kind, ctx.cg_iters(), w.cg_iters())
            up, np2, cum = ctx.download_particles()
            assert np.array_equal(cum, w.array(rank, O.CUMCNT)), "%s: per-cell counts differ on rank %d step %d" % (kind, rank, it)
            a, b = flatten_by_id(up, np2), flatten_by_id(w.array(rank, O.UP), w.array(rank, O.NP2))
            assert np.array_equal(a[0], b[0])
            ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
            tol = 1e-12 if it == 0 else 1e-10
            assert ex <= tol and eu <= tol, (kind, it, ex, eu)
            assert rel_to_max(ctx.download_field(), w.array(rank, O.UF)).max() <= tol, kind
        ctx.close()
        w.close()
    dist.barrier()
    dist.destroy_process_group()
    print("rank %d/%d ok" % (rank, world))


if __name__ == "__main__":
    main()
