#!/usr/bin/env python
"""Parity against the REAL reference binary, for whoever has gfortran + MPI (this image has neither).

What the maintainer produces with the unmodified reference (any rank count, e.g. proj/weibel with config_sample.json):

    1. run to some step N and let it write   NNNNNNN_restart.json + .raw          (save_restart, proj/weibel/app.f90:550-580;
       e.g. max_elapsed small, or max_it = N)
    2. restart from that file ("restart_file" in config.json), run K more steps (K = 1 is the sharpest test) and write
       MMMMMMM_restart.json + .raw

    python scripts/compare_with_reference_run.py  NNNNNNN_restart  MMMMMMM_restart  [--gfac 0.501] [--bc periodic|reconnection]

The script loads snapshot 1 into ONE device context that holds all rows (the ranks' slabs concatenated along y: the
decomposition must not matter, tests/config1_worker.py checks that), re-buckets the rows exactly like the reference's restart
path (io__input + sort__bucket, proj/weibel/app.f90:349-353; the CG warm start df starts from zero on both sides: it is a
SAVE variable that no snapshot holds, common/field.f90:98), advances K = it2 - it1 steps and compares with snapshot 2:

    per-cell counts            bit-exact
    particles matched by ID    <= 1e-12 (positions relative to nx*delx, momenta relative to max(|u|, 1e-3 c))   for K = 1,
                               growing slowly with K (trajectories decorrelate chaotically; use K <= 10)
    uf                         <= 1e-12 relative to each component's max-abs for K = 1

Exit code 0 = parity with the Fortran binary holds; the numbers are printed either way.  This is the one file pair that
turns "parity unpinned"  into a pin.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def stack_ranks(up, np2, uf):
    """per-rank arrays (snapshot.read_restart) -> one slab holding all rows"""
    upa = np.concatenate(up, axis=1)                      # (nsp, nyl * nproc, np, 6)
    np2a = np.concatenate(np2, axis=1)
    if len(uf) == 1:
        ufa = uf[0]
    else:
        ufa = np.concatenate([uf[0][:-2]] + [u[2:-2] for u in uf[1:-1]] + [uf[-1][2:]], axis=0)
    return np.ascontiguousarray(upa), np.ascontiguousarray(np2a), np.ascontiguousarray(ufa)


def by_id(up, np2):
    from helpers import flatten_by_id
    return flatten_by_id(up, np2)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("snap1")
    ap.add_argument("snap2")
    ap.add_argument("--gfac", type=float, default=0.501)           # proj/*/app.f90:64-68, not part of a snapshot
    ap.add_argument("--bc", default="periodic", choices=["periodic", "reconnection"])
    ap.add_argument("--tol", type=float, default=1e-12)
    args = ap.parse_args()
    import wumingpic2d_b200 as wm
    from wumingpic2d_b200 import snapshot as S
    a1, up1, np21, uf1 = S.read_restart(args.snap1)
    a2, up2, np22, uf2 = S.read_restart(args.snap2)
    for k in ("nxgs", "nxge", "nygs", "nyge", "nsp", "np", "delx", "delt", "c"):
        assert a1[k] == a2[k], "the two snapshots are not of the same run (%s: %r vs %r)" % (k, a1[k], a2[k])
    steps = a2["it"] - a1["it"]
    assert steps >= 1, "snapshot 2 must be later than snapshot 1"
    up, np2, uf = stack_ranks(up1, np21, uf1)
    upr, np2r, ufr = stack_ranks(up2, np22, uf2)
    nx, ny = a1["nxge"] - a1["nxgs"] + 1, a1["nyge"] - a1["nygs"] + 1
    ctx = wm.Context(np_cap=a1["np"], nxgs=a1["nxgs"], nxge=a1["nxge"], nygs=a1["nygs"], nyge=a1["nyge"], nys=a1["nygs"],
                     nye=a1["nyge"], delx=a1["delx"], delt=a1["delt"], c=a1["c"], q=list(np.atleast_1d(a1["q"])),
                     r=list(np.atleast_1d(a1["r"])), gfac=args.gfac, nsp=a1["nsp"],
                     bc=wm.WM_BC_PERIODIC if args.bc == "periodic" else wm.WM_BC_RECONNECTION)
    ctx.upload_particles(up, np2)            # rows are re-bucketed: io__input + sort__bucket
    ctx.upload_field(uf)
    ctx.step(steps)
    upg, np2g, cumg = ctx.download_particles()
    ufg = ctx.download_field()
    # the reference's snapshot rows are lists in arrival order: bucket-sort them for the per-cell counts
    ok = True
    cnt_ref = np.zeros((a1["nsp"], ny, nx), np.int64)
    for isp in range(a1["nsp"]):
        for j in range(ny):
            xs = upr[isp, j, :np2r[isp, j], 0].astype(np.int64) - a1["nxgs"]
            cnt_ref[isp, j] = np.bincount(xs, minlength=nx)[:nx]
    same_counts = np.array_equal(np.diff(cumg, axis=2), cnt_ref) and np.array_equal(np2g, np2r)
    print("steps %d   per-cell counts bit-exact: %s" % (steps, same_counts))
    ok &= bool(same_counts)
    ga, gb = by_id(upg, np2g), by_id(upr, np2r)
    if not (np.array_equal(ga[0], gb[0]) and np.array_equal(ga[1], gb[1])):
        print("particle id sets differ")
        ok = False
    else:
        ex = np.abs(ga[2][:, :2] - gb[2][:, :2]).max() / (nx * a1["delx"])
        eu = (np.abs(ga[2][:, 2:] - gb[2][:, 2:]) / np.maximum(np.abs(gb[2][:, 2:]), 1e-3 * a1["c"])).max()
        print("particles by id: position error %.3e (of nx*delx), momentum error %.3e (relative)" % (ex, eu))
        ok &= bool(ex <= args.tol * max(1, steps) and eu <= args.tol * max(1, steps) * 10)
    d = np.abs(ufg[2:-2, 2:-2] - ufr[2:-2, 2:-2]).reshape(-1, 6).max(axis=0)
    sc = np.abs(ufr[2:-2, 2:-2]).reshape(-1, 6).max(axis=0)
    ef = (d / np.where(sc > 0, sc, 1.0)).max()
    print("uf interior: max error %.3e relative to each component's max-abs; CG iterations on the device %s" % (ef, ctx.cg_iters()))
    ok &= bool(ef <= args.tol * max(1, steps) * 10)
    print("PARITY WITH THE REFERENCE BINARY: %s" % ("ok" if ok else "FAILED"))
    ctx.close()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
