#!/usr/bin/env python
"""bench.py -- particle-steps/s of the per-timestep PIC hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], SURVEY 8d "Config 2"): Weibel set-up, synthetic uniform
Maxwellian, nx = 4096, 512 rows per GPU (N GPUs -> 4096 x 512N grid; 8 GPUs = 4096 x 4096),
64 particles per cell per species, e- + ion: 268,435,456 particles per GPU.  Weak scaling.

A step = push + Esirkepov deposit + particle boundaries + field solve (CG) + migration + sort,
i.e. the five calls of proj/weibel/app.f90:100-107, fused (wm_step).

value   particle-steps/s, state resident in HBM, K steps timed with CUDA events on the library's
        stream inside wm_step (first launch -> last completion), max over ranks.
e2e     the same metric through wm_host_step: HOST arrays in the reference's layout (pinned), H2D of
        up/uf/np2/cumcnt and D2H of the same inside the timed region, every step.
roofline the fused push+deposit+boundary kernel: algorithmic 96 B per particle (48 B record read
        once, written once) / its event-timed duration, vs the measured HBM copy peak.
cpu_baseline the CPU oracle (a C++/OpenMP restatement of the reference path, NOT the Fortran
        build, which cannot be produced here: no Fortran compiler, no MPI) on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NX, ROWS_PER_GPU, PPC = 4096, 512, 64
SEED = 20260117
ALG_BYTES_STEP = 192.0    # 4 x 48 B per particle-step
ALG_BYTES_PASS1 = 96.0    # fused push+deposit+boundary pass: R 48 + W 48


def weibel_params(nx, ny, n_ppc, cap_factor=1.25):
    """proj/weibel/app.f90:245-255,292-303 with proj/weibel/config_sample.json physics."""
    import math
    c, gfac, cfl, delx = 1.0, 0.501, 1.0, 1.0
    mass_ratio, sigma_e, omega_pe, v_the, v_thi, t_ani = 1.0, 0.0, 0.1, 0.1, 0.1, 5.0
    delt = cfl * delx / c
    wpe = omega_pe
    wge = omega_pe * math.sqrt(sigma_e)
    wpi = wpe / math.sqrt(mass_ratio)
    wgi = wge / mass_ratio
    r = [mass_ratio, 1.0]
    q = [+math.sqrt(r[0] / (4 * math.pi * n_ppc / delx ** 2)) * wpi,
         -math.sqrt(r[1] / (4 * math.pi * n_ppc / delx ** 2)) * wpe]
    return dict(nx=nx, ny=ny, nranks=1, n0=n_ppc, np=int(math.ceil(n_ppc * nx * cap_factor)), nsp=2,
                delx=delx, delt=delt, c=c, gfac=gfac, q=q, r=r, b0=r[0] * c / q[0] * wgi,
                vti=v_thi, vte=v_the, t_ani=t_ani, nxgs=2, nygs=2)


class ClockSampler:
    """SM clock and throttle reasons of this rank's GPU DURING the timed region (B200_PROFILING.md): NVML polled every 5 ms by a
    thread of this process (nvidia-smi needs longer to start on an 8-GPU box than 20 steps take); `nvidia-smi -lms` as the
    fallback when NVML cannot be loaded."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.lines, self.p, self.h, self.nv = index, [], None, None, None
        self.sm, self.mask, self.mx, self.stop_flag = [], 0, None, False
        try:  # the handle is opened before the timed region; by UUID, so that CUDA_VISIBLE_DEVICES does not matter
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(index).uuid))
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.h is not None:
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.h is not None:
            self.stop_flag = True
            self.t.join(timeout=2)
            sm = sorted(self.sm)
            return {"sm_mhz": (sm[len(sm) // 2] if sm else None), "sm_max_mhz": self.mx, "samples": len(sm),
                    "reasons": [n for b, n in self.REASONS if self.mask & b], "source": "nvml, every 5 ms inside the timed region"}
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # under load = upper half of the samples (the region is short; idle samples bracket it)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


def workload_config(nx, rows, ppc, world):
    """The `config` object of both arms (the reference arm runs a bounded sample of exactly this workload)."""
    return {"workload": "weibel %dx%d grid (%d rows per GPU), %d ppc/species, e-/ion, uniform Maxwellian "
                        "(BASELINE configs[1] slab)" % (nx, rows * world, rows, ppc),
            "particles": 2 * nx * rows * world * ppc, "parallelism": "y-slab x%d" % world}


def host_threads():
    """All host cores of the box.  Set explicitly: torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, and
    under it only rank 0 runs the CPU arm, so it takes every core (the ranks do not share the CPU arm)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bind_to_gpu_numa(local):
    """Pin this rank to the CPUs next to its GPU before the pinned host buffers of the e2e leg are allocated (first touch
    decides the NUMA node): with 8 ranks each moving 26 GB per step through host memory, buffers on the far socket halve the
    PCIe rate.  Returns the CPU list, or None if the topology cannot be read."""
    try:
        bus = subprocess.check_output(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                                      text=True, stderr=subprocess.DEVNULL).strip().lower()
        dom, rest = bus.split(":", 1)
        with open("/sys/bus/pci/devices/%s:%s/local_cpulist" % (dom[-4:], rest)) as f:
            txt = f.read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return txt
    except Exception:
        pass
    return None


def cpu_reference_rate(steps, warmup, rows=32, nx=NX, ppc=PPC):
    """The CPU oracle (timing build, all host threads) on a bounded sample of the workload:
    same nx, ppc, physics and IC recipe, `rows` rows instead of 512 per GPU."""
    import oracle_lib as O
    prm = weibel_params(nx, rows, ppc)
    w = O.World(prm, fast=True)
    w.lib.orc_set_num_threads(host_threads())
    w.ic_weibel(SEED)
    npart = 2 * nx * rows * ppc
    if warmup:
        w.step(warmup)
    w.stage_times(reset=True)
    t0 = time.perf_counter()
    w.step(steps)
    dt = time.perf_counter() - t0
    st = w.stage_times()
    cores = w.lib.orc_num_threads()
    w.close()
    return dict(value=npart * steps / dt, unit="particle-steps/s", cores=cores, kind="port",
                sample="%dx%d grid, %d ppc/species x2 (%d particles), %d steps after %d warm-up; "
                       "C++/OpenMP restatement of the reference CPU path (Fortran+MPI build impossible here)"
                       % (nx, rows, ppc, npart, steps, warmup),
                ms_per_step=1e3 * dt / steps,
                stage_s=dict(zip(("push", "field+deposit", "bc_x", "bc_y", "sort"), [round(x, 4) for x in st])))


def run_reference(args, rank, world):
    """The reference's CPU path (C++/OpenMP restatement: the Fortran + MPI build cannot be produced here) on all host
    cores of the box, rank 0 only, on a bounded sample of the GPU arm's workload: the same nx, ppc, physics and initial
    condition, REF_ROWS rows (4096 x 256 cells x 64 ppc x 2 = 134 M particles -- the particle count of BASELINE.md
    section 4's 1024^2 sample, kept at the slab's own nx so the row length / cache behaviour is the workload's).  The
    metric is a rate, so it does not depend on the number of rows once the working set is far beyond the caches."""
    if rank != 0:
        return
    base = cpu_reference_rate(args.steps, args.warmup, rows=args.ref_rows, nx=args.nx, ppc=args.ppc)
    line = {
        "impl": "reference", "metric": "particle-steps/sec (push+deposit+sort+field)", "value": base["value"],
        "unit": "particle-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.nx, args.rows, args.ppc, args.gpus),
        "cpu_parallelism": "openmp%d (rank 0 of %d, all host cores)" % (base["cores"], world),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--rows", type=int, default=ROWS_PER_GPU, help="rows per GPU (default: the named workload)")
    ap.add_argument("--nx", type=int, default=NX)
    ap.add_argument("--ppc", type=int, default=PPC)
    ap.add_argument("--ref-rows", type=int, default=256, help="rows of the CPU arm's sample (--impl reference)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-interval", type=int, default=50, help="also time wm_host_steps over this many steps per upload/download")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-affinity", action="store_true", help="N > 1: do not pin the ranks to the CPUs next to their GPUs")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--exact", action="store_true", help="bit-exact push arithmetic (WM_FLAG_EXACT_PUSH)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            # convenience: re-launch under torchrun
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                   "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
            sys.exit(subprocess.call(cmd))
        raise SystemExit("WORLD_SIZE=%d but --gpus %d" % (world, args.gpus))

    import numpy as np
    import torch
    import wumingpic2d_b200 as wm

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    numa = bind_to_gpu_numa(local) if (world > 1 and not args.no_affinity) else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nx, rows, ppc = args.nx, args.rows, args.ppc
    ny = rows * world
    prm = weibel_params(nx, ny, ppc)
    nys = 2 + rank * rows
    nye = nys + rows - 1
    ctx = wm.Context.from_params(prm, nys=nys, nye=nye, nrank=rank, nsize=world, device=local,
                                 flags=wm.WM_FLAG_EXACT_PUSH if args.exact else 0)
    if world > 1:
        idbuf = [ctx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(idbuf, src=0)
        ctx.comm_init(idbuf[0])
    ctx.ic_weibel(SEED, ppc, prm["vti"], prm["vte"], prm["t_ani"], prm["b0"])
    n_local = sum(ctx.particle_counts())
    n_total = allsum(float(n_local))

    # ---- resident throughput ---------------------------------------------------------------
    ctx.step(args.warmup)
    ctx.timing(reset=True)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    ctx.step(args.steps)
    ctx.synchronize()
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    ms, launches = ctx.timing(reset=True)
    ms_step = allmax(ms[4]) / args.steps
    wall_step = allmax(wall) * 1e3 / args.steps
    value = n_total / (ms_step * 1e-3)
    ms_pass1 = allmax(ms[0]) / args.steps
    cg = ctx.cg_iters()

    # ---- did the physics survive?  (every N: particle number over all ranks, discrete Gauss law over the ring)
    n_after = allsum(float(sum(ctx.particle_counts())))
    g_res, g_scale = ctx.gauss_residual()
    g_res, g_scale = allmax(g_res), allmax(g_scale)
    check = {"particles_conserved": bool(n_after == n_total), "particles": int(n_after),
             "gauss_rel": g_res / g_scale if g_scale > 0 else None, "gauss_tol": 1e-10,
             "steps_run": args.warmup + args.steps,
             "what": "sum over ranks of the particle counts after the run == before; max |div E - 4 pi rho| / max 4 pi rho_abs "
                     "over all cells of all ranks (rho folded over the rank ring), must stay at roundoff (Esirkepov)"}
    check["ok"] = bool(check["particles_conserved"] and check["gauss_rel"] is not None and check["gauss_rel"] <= check["gauss_tol"])

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_kind = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    achieved = ALG_BYTES_PASS1 * n_local / (ms_pass1 * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum and the FP64 pipe utilisation of the same kernel, from the committed
    # `ncu --set full` capture of this workload (profiles/pass1_traffic.json names the report it was read from)
    traffic, fp64_pipe = None, None
    # FP64 vector peak of this device, measured in this run with a DFMA loop; the kernel executes 285
    # FP64 instructions per particle over the whole launch (DFMA 180 + DMUL 61 + DADD 31 + DSETP 13: executed-instruction
    # counts of the ncu capture, profiles/r02_fused_sm.md; 209 in the particle loop alone), so its FP64 instruction rate is known live
    try:
        fp64_peak = ctx.fp64_peak()
    except Exception:
        fp64_peak = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "pass1_traffic.json")))
        if (nx, rows, ppc) == (tj.get("nx"), tj.get("rows"), tj.get("ppc")) and not args.exact:
            traffic = tj.get("dram_bytes_per_launch")
        fp64_pipe = tj.get("fp64_pipe_pct_of_peak")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "k_pass1<PUSH|DEPOSIT|BOUND> (exact)" if args.exact else {"5": "k_fused_dp (push + Esirkepov deposit split into stayers/movers + particle boundaries + cell sort with direct placement: every record written once into the other store; one pass)", "1": "k_fused_sm<TAIL> (push + Esirkepov deposit split into stayers/movers + particle boundaries + in-place cell sort: stayers compacted, in-tile cell changers appended to their new segments by the CTA's tail; one pass)", "0": "k_fused<INPLACE>"}.get(os.environ.get("WM_SM", "1"), "k_fused_sm variant"),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_kind, "alg_bytes_per_particle": ALG_BYTES_PASS1, "ms_per_launch": ms_pass1,
                "fp64_pipe_pct_of_peak_ncu": fp64_pipe,
                "fp64_peak_tflops_measured": fp64_peak,
                "fp64_instr_per_particle": 285.0,
                "fp64_instr_rate_frac_of_measured_peak": (285.0 * n_local / (ms_pass1 * 1e-3) * 2 / 1e12 / fp64_peak) if fp64_peak else None,
                "note": ("alg_bytes_per_particle = 48 B read + 48 B written per particle (SURVEY 8d); the kernel also does the sort "
                         "of 98 % of the particles (stayers compacted in place, in-tile cell changers staged and appended by the "
                         "CTA's tail: about 13 B/particle more traffic, not counted), which SURVEY 8d books as another 96 B: "
                         "compare whole_step across rounds as well"),
                "whole_step": {"alg_bytes_per_particle_step": ALG_BYTES_STEP,
                               "achieved": ALG_BYTES_STEP * n_local / (allmax(ms[4]) / args.steps * 1e-3) / 1e9,
                               "frac": ALG_BYTES_STEP * n_local / (allmax(ms[4]) / args.steps * 1e-3) / 1e9 / peak}}
    stage_ms = {"fused_push_deposit_boundary_sort": ms[0] / args.steps, "field_solve": ms[1] / args.steps,
                "prep_and_migration": ms[2] / args.steps, "sort_place_cell_changers": ms[3] / args.steps}

    # ---- end to end through host arrays -------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        try:
            shape_up = ctx.shape_up()
            nbytes_up = int(np.prod(shape_up)) * 8
            up_t = torch.empty(nbytes_up // 8, dtype=torch.float64, pin_memory=True)
            uf_t = torch.empty(int(np.prod(ctx.shape_uf())), dtype=torch.float64, pin_memory=True)
            up = up_t.numpy().reshape(shape_up)
            uf = uf_t.numpy().reshape(ctx.shape_uf())
            _, np2_d, cum_d = ctx.download_particles(up)
            # the index arrays are pinned as well: a copy from or to pageable memory blocks the host thread that feeds the pipeline
            np2_t = torch.empty(np2_d.size, dtype=torch.int32, pin_memory=True)
            cum_t = torch.empty(cum_d.size, dtype=torch.int32, pin_memory=True)
            np2, cum = np2_t.numpy().reshape(np2_d.shape), cum_t.numpy().reshape(cum_d.shape)
            np2[...], cum[...] = np2_d, cum_d
            del np2_d, cum_d
            uf[...] = ctx.download_field()
            h2d = d2h = int(np2.sum()) * 48 + uf.nbytes + np2.nbytes + cum.nbytes
            ctx.host_step(up, uf, np2, cum)  # warm-up
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.e2e_steps):
                ctx.host_step(up, uf, np2, cum)
            barrier()
            dt = allmax(time.perf_counter() - t0)
            e2e = {"value": n_total * args.e2e_steps / dt, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps,
                   "api": "wm_host_step(up, uf, np2, cumcnt): pinned host arrays in the reference's Fortran layout",
                   "cpu_affinity_rank0": numa,
                   "host_pipe_chunks": ctx.host_pipe_chunks(),
                   "note": "every step moves the whole state over PCIe, %.1f GB up and %.1f GB down; the rows travel in chunks of "
                           "WM_HOSTPIPE_ROWS (default 16) rows, upload of the next chunks / particle pass / download of the finished "
                           "rows side by side: bounded by ~13 GB / 45 GB/s of full-duplex PCIe Gen5 (measured, "
                           "scripts/micro/pcie_duplex.py) instead of 2 x 13 GB / 55 GB/s" % (h2d / 1e9, d2h / 1e9)}
            # how the shim is meant to be used: host arrays refreshed every intvl_mom = 50 steps (WM_SYNC_INTERVAL=50)
            if args.e2e_interval > 1:
                barrier()
                t0 = time.perf_counter()
                ctx.host_steps(up, uf, np2, cum, args.e2e_interval)
                barrier()
                dti = allmax(time.perf_counter() - t0)
                e2e["sync_interval_%d" % args.e2e_interval] = {
                    "value": n_total * args.e2e_interval / dti, "unit": "particle-steps/s", "ms_per_step": 1e3 * dti / args.e2e_interval,
                    "h2d_bytes_per_step": h2d // args.e2e_interval, "d2h_bytes_per_step": d2h // args.e2e_interval,
                    "api": "wm_host_steps(up, uf, np2, cumcnt, %d): one upload, %d steps, one download (the shim's "
                           "WM_SYNC_INTERVAL = intvl_mom of proj/weibel/config_sample.json)" % (args.e2e_interval, args.e2e_interval)}
            del up, uf, up_t, uf_t
        except Exception as ex:  # keep the resident number even if host memory is short
            e2e = {"value": None, "unit": "particle-steps/s", "error": str(ex)[:200]}

    # ---- CPU baseline (rank 0, N = 1 only) ----------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            b = cpu_reference_rate(steps=10, warmup=2)
            cpu = {k: b[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as ex:
            cpu = {"value": None, "error": str(ex)[:200]}

    if rank == 0:
        line = {
            "metric": "particle-steps/sec (push+deposit+sort+field)", "value": value, "unit": "particle-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(nx, rows, ppc, world),
            "run_info": {"l2": "working set %.1f GB per GPU >> 126 MB L2: no flush needed between steps" % (n_local * 96 / 1e9),
                         "push_arithmetic": "exact" if args.exact else "fma", "cg_iters": cg,
                         "cg_path": {0: "host loop of small kernels (+ NCCL per iteration on a ring)",
                                     1: "persistent cooperative kernel", 2: "persistent cooperative kernel, ring exchange and "
                                     "all-reduce in the kernel over CUDA-IPC peer memory"}[ctx.cg_path()],
                         "sort": "tag+scatter" if (args.exact or os.environ.get("WM_INPLACE") == "0") else "in-place",
                         "layout_rebuilds": ctx.rebuilds()},
            "check": check,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": clocks, "stage_ms": stage_ms, "wall_ms_per_step": wall_step,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    if not check["ok"]:
        raise SystemExit("bench.py: correctness check failed: %s" % json.dumps(check))


if __name__ == "__main__":
    main()
