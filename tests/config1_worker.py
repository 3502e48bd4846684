"""BASELINE configs[0]: proj/weibel/config_sample.json at its real size -- 256 x 256 cells, 20 particles per cell and
species, num_process = 4 (64 rows per slab), 1000 steps, energy every intvl_mom = 50 steps -- on WORLD_SIZE GPUs against the
oracle's world with the same number of slabs (common/mpi_set.f90:36-47).  Launched under torch.distributed.run by
tests/test_gpu_multi.py (4 GPUs) or run stand-alone on one GPU (one slab of 256 rows against the oracle's 4-slab world: the
decomposition must not matter).

Checks: after the first step per-cell counts bit-exact, particles by ID and fields <= 1e-12, equal CG iteration counts; every
50 steps the total energy (all-reduced over the ranks) within 1e-6 relative of the oracle's, every component within 1e-3 of
the total; particle number conserved; the Weibel instability has grown.  Prints one JSON line (rank 0) that is kept under
profiles/.  Exit code 0 = all checks hold."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

import oracle_lib as O  # noqa: E402
from helpers import flatten_by_id, particle_err, rel_to_max  # noqa: E402

# proj/weibel/config_sample.json ("parameter" section) and the constants of proj/weibel/app.f90:64-68
CONFIG1 = dict(n_x=256, n_y=256, n_ppc=20, num_process=4, mass_ratio=1.0, sigma_e=0.0, omega_pe=0.1, v_the=0.1, v_thi=0.1,
               t_ani=5.0, max_it=1000, intvl_mom=50)


def main():
    import torch
    import wumingpic2d_b200 as wm
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else CONFIG1["max_it"]
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = CONFIG1
    oranks = world if world > 1 else g["num_process"]
    prm = O.weibel_params(g["n_x"], g["n_y"], g["n_ppc"], nranks=oranks, mass_ratio=g["mass_ratio"], sigma_e=g["sigma_e"],
                          omega_pe=g["omega_pe"], v_the=g["v_the"], v_thi=g["v_thi"], t_ani=g["t_ani"])
    w = O.World(prm, fast=False)
    w.lib.orc_set_num_threads(max(1, (os.cpu_count() or 8) // world))
    w.ic_weibel(20260117)
    if world > 1:
        nys, nye = w.bounds(rank)
        ctx = wm.Context.from_params(prm, nys=nys, nye=nye, nrank=rank, nsize=world, device=local)
        ids = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0])
        mine = [rank]
    else:
        prm1 = dict(prm, nranks=1)
        ctx = wm.Context.from_params(prm1, device=local)
        mine = list(range(oranks))

    def ostate(which):
        return np.concatenate([w.array(r, which) for r in mine], axis=1 if which in (O.UP, O.NP2, O.CUMCNT) else 0)

    def ofield():
        parts = [w.array(r, O.UF) for r in mine]
        if len(parts) == 1:
            return parts[0]
        return np.concatenate([parts[0][:-2]] + [p[2:-2] for p in parts[1:-1]] + [parts[-1][2:]], axis=0)

    ctx.upload_particles_sorted(np.ascontiguousarray(ostate(O.UP)), np.ascontiguousarray(ostate(O.NP2)), np.ascontiguousarray(ostate(O.CUMCNT)))
    ctx.upload_field(np.ascontiguousarray(ofield()))

    def allsum(v):
        if dist is None:
            return v
        t = torch.tensor(v, dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return t.cpu().numpy()

    e0 = w.energy()
    eg0 = allsum(ctx.energy())
    assert abs(eg0.sum() - e0.sum()) <= 1e-12 * e0.sum()
    # ---- one step: full parity
    w.step(1)
    ctx.step(1)
    assert ctx.cg_iters() == w.cg_iters(), (ctx.cg_iters(), w.cg_iters())
    up, np2, cum = ctx.download_particles()
    assert np.array_equal(cum, ostate(O.CUMCNT)), "per-cell counts differ after one step (rank %d)" % rank
    a, b = flatten_by_id(up, np2), flatten_by_id(ostate(O.UP), ostate(O.NP2))
    assert np.array_equal(a[0], b[0])
    ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
    ef = float(rel_to_max(ctx.download_field()[2:-2], ofield()[2:-2]).max())
    assert ex <= 1e-12 and eu <= 1e-12 and ef <= 1e-12, (ex, eu, ef)
    # ---- the energy history
    hist, worst, t_gpu, done = [], 0.0, 0.0, 1
    while done < nsteps:
        n = min(g["intvl_mom"] - done % g["intvl_mom"], nsteps - done)
        w.step(n)
        t0 = time.perf_counter()
        ctx.step(n)
        ctx.synchronize()
        t_gpu += time.perf_counter() - t0
        done += n
        eo, eg = w.energy(), allsum(ctx.energy())
        rel = abs(eg.sum() - eo.sum()) / eo.sum()
        worst = max(worst, rel)
        hist.append([done, float(eo.sum()), float(eg.sum())])
        assert rel <= 1e-6, "step %d: total energy GPU %r vs oracle %r" % (done, eg.sum(), eo.sum())
        assert np.abs(eg - eo).max() <= 1e-3 * eo.sum(), "step %d: an energy component drifted" % done
    ntot = allsum(np.array([float(sum(ctx.particle_counts()))]))[0]
    assert ntot == 2.0 * g["n_x"] * g["n_y"] * g["n_ppc"]
    if nsteps >= 500:
        assert eg[-1] > 1e-4 * e0.sum(), "the Weibel instability should have grown by now"
    drift_o, drift_g = eo.sum() - e0.sum(), eg.sum() - e0.sum()
    assert abs(drift_g - drift_o) <= 0.1 * abs(drift_o) + 1e-9 * e0.sum()
    if rank == 0:
        print(json.dumps({"config": "proj/weibel/config_sample.json (256x256, 20 ppc, 2 species)", "gpus": world,
                          "oracle_slabs": oranks, "steps": done, "first_step": {"pos_err": ex, "mom_err": eu, "field_err": ef,
                          "cg_iters": ctx.cg_iters(), "counts_bit_exact": True},
                          "energy_rel_err_max": worst, "energy_tol": 1e-6, "energy_history_every_50": hist[-3:],
                          "gpu_ms_per_step": 1e3 * t_gpu / max(done - 1, 1), "cg_path": ctx.cg_path(),
                          "layout_rebuilds": ctx.rebuilds()}), flush=True)
    ctx.close()
    w.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    print("rank %d/%d ok" % (rank, world), flush=True)


if __name__ == "__main__":
    main()
