#!/usr/bin/env python
"""BASELINE configs[2] and [3] at scale on the device -- the non-uniform loads:

    harris : proj/reconnection Harris current sheet, 2048 x (512 per GPU) cells, nbg = 100, ncs = 500 (SURVEY 8d "Config 3")
    shock  : proj/shock, 16384 x (128 per GPU) cells, box starts at n_x_ini and grows one column per step, inject + relocate
             every step on the device (SURVEY 8d "Config 4")

    python scripts/run_configs.py harris|shock [--steps K] [--warmup W] [--small]
    python -m torch.distributed.run --nproc-per-node N ... scripts/run_configs.py harris --gpus N

Prints one JSON line (rank 0): ms per step (CUDA events inside wm_step, max over ranks), particle-steps/s, layout rebuilds,
the static-tile cost of the load (lane efficiency of k_fused_sm's 8-lanes-per-cell mapping from the per-cell counts, the
spread of the per-tile trip counts) and the correctness bits (particle number, energy drift).  The lines are kept under
profiles/.
"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def tile_stats(cum, tx=16, ty=8):
    """Work of the static tile grid: a warp of k_fused_sm takes 4 cells adjacent in x, 8 lanes each, and runs
    max over the 4 cells of ceil(count / 8) iterations per species; a CTA is one 16 x 8 tile."""
    import numpy as np
    cnt = np.diff(cum, axis=2)                       # (nsp, nyl, nx)
    nsp, nyl, nx = cnt.shape
    nxp, nyp = -(-nx // tx) * tx, -(-nyl // ty) * ty
    pad = np.zeros((nsp, nyp, nxp), dtype=np.int64)
    pad[:, :nyl, :nx] = cnt
    it = -(-pad // 8)
    quad = it.reshape(nsp, nyp, nxp // 4, 4).max(axis=3)            # iterations of a warp per (species, quad)
    trips = quad.sum(axis=0)                                          # both species
    tile = trips.reshape(nyp // ty, ty, nxp // tx, tx // 4).sum(axis=(1, 3))   # warp-iterations per tile (4 warps share them)
    tot = float(cnt.sum())
    return {"lane_efficiency": tot / (32.0 * trips.sum()) if trips.sum() else None,
            "tile_trips_min": int(tile.min()), "tile_trips_median": float(np.median(tile)), "tile_trips_max": int(tile.max()),
            "tile_trips_mean": float(tile.mean()), "empty_tiles_frac": float((tile == 0).mean()),
            "cell_count_max": int(cnt.max()), "cell_count_mean": float(cnt.mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kind", choices=["harris", "shock"])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--small", action="store_true", help="1/8 of the slab (quick check)")
    ap.add_argument("--ppc", type=int, default=32, help="shock: particles per cell and species upstream")
    ap.add_argument("--capf", type=float, default=8.0, help="shock: device slots per species = capf x n0 x nx x rows")
    args = ap.parse_args()
    import numpy as np
    import torch
    import wumingpic2d_b200 as wm
    from helpers import harris_params, shock_params
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def allred(x, op="sum"):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    if args.kind == "harris":
        nx, rows = (2048, 512) if not args.small else (512, 128)
        nbg, ncs = 100, 500
        prm = harris_params(nx, rows * world, nbg, ncs, nranks=world)
        nys = 2 + rank * rows
        ctx = wm.Context.from_params(prm, nys=nys, nye=nys + rows - 1, nrank=rank, nsize=world, device=local, bc=wm.WM_BC_RECONNECTION)
    else:
        nx, rows = (16384, 128) if not args.small else (2048, 64)
        n0 = args.ppc
        prm = shock_params(nx, rows * world, n0, nranks=world, u_inject=40.0, sigma_e=0.1, v_the=0.01, v_thi=0.01, l_damp_ini=100.0)
        nys = 2 + rank * rows
        ctx = wm.Context.from_params(prm, nys=nys, nye=nys + rows - 1, nrank=rank, nsize=world, device=local, bc=wm.WM_BC_SHOCK,
                                     capacity=int(n0 * nx * rows * args.capf))
    if world > 1:
        ids = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.comm_init(ids[0])
    seed = 20260117
    t0 = time.perf_counter()
    if args.kind == "harris":
        ctx.ic_harris(seed, nbg, ncs, prm["lcs_cells"], prm["vti"], prm["vte"], prm["b0"], prm["rtemp"], 0.12)
    else:
        nxe0 = prm["nxgs"] + nx // 2                           # n_x_ini = n_x / 2 as in proj/shock/config_sample.json
        ctx.ic_shock(seed, n0, nxe0, prm["v0"], prm["vti"], prm["vte"], prm["b0"], prm["theta"], prm["phi"], prm["l_damp"])
        ctx.set_u_inject(prm["u0"])
    ctx.synchronize()
    t_ic = time.perf_counter() - t0
    n_start = sum(ctx.particle_counts())
    e_start = ctx.energy()

    def advance(n, it0):
        """n steps of the application's loop: proj/reconnection/app.f90:100-107, proj/shock/app.f90:109-125"""
        src = 0.0
        if args.kind == "harris":
            ctx.step(n)
            return src
        for k in range(n):
            ctx.step(1)
            t1 = time.perf_counter()
            ctx.shock_inject(seed, it0 + k + 1)
            ctx.shock_relocate(seed, it0 + k + 1)                # intvl_expand = 1
            ctx.synchronize()
            src += time.perf_counter() - t1
        return src

    advance(args.warmup, 0)
    reb0 = ctx.rebuilds()
    ctx.timing(reset=True)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    tw = time.perf_counter()
    t_src = advance(args.steps, args.warmup)
    ctx.synchronize()
    wall = time.perf_counter() - tw
    ms, launches = ctx.timing(reset=True)
    ms_step = allred(ms[4], "max") / args.steps
    n_end = sum(ctx.particle_counts())
    e_end = ctx.energy()
    _, np2, cum = ctx.download_particles(want_up=False)
    ts = tile_stats(cum)
    n_tot = allred(float(n_end))
    line = {
        "config": {"harris": "proj/reconnection Harris sheet, %dx%d grid (%d rows per GPU), nbg=100 ncs=500, mass ratio 16, cfl 0.5 (BASELINE configs[2])",
                   "shock": "proj/shock, %dx%d grid (%d rows per GPU), n_x_ini = n_x/2, u_inject 40, inject + relocate every step on the device (BASELINE configs[3])"}[args.kind]
                  % (nx, rows * world, rows),
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "particles": int(n_tot),
        "ms_per_step_device": ms_step, "particle_steps_per_s": n_tot / (ms_step * 1e-3),
        "wall_ms_per_step": 1e3 * allred(wall, "max") / args.steps,
        "stage_ms": {"fused": ms[0] / args.steps, "field": ms[1] / args.steps, "prep_migration": ms[2] / args.steps, "sort_tail": ms[3] / args.steps},
        "layout_rebuilds_in_timed_steps": int(allred(float(ctx.rebuilds() - reb0), "max")),
        "static_tile_grid": ts, "cg_iters": ctx.cg_iters(), "cg_path": ctx.cg_path(),
        "ic_seconds": t_ic, "particles_rank0": [int(n_start), int(n_end)],
        "energy_rank0": {"start": float(e_start.sum()), "end": float(e_end.sum())},
        "uniform_reference": "uniform 64 ppc Weibel slab: 12.7 ms per 268 M particles = 21.1 G particle-steps/s per GPU",
    }
    if args.kind == "shock":
        line["sources_ms_per_step_host_wall"] = 1e3 * t_src / args.steps
        line["xrange_end"] = list(ctx.xrange())
    if args.kind == "harris":
        line["check"] = {"particles_conserved": n_start == n_end,
                         "energy_rel_change": abs(float(e_end.sum() - e_start.sum())) / float(e_start.sum())}
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
