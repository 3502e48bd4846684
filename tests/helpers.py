"""Shared helpers for the parity tests (CUDA path through the C ABI vs the CPU oracle)."""
import numpy as np

import oracle_lib as O


def flatten_by_id(up, np2):
    """(nsp, nyl, np, 6) host array + np2 -> ids, species, rec[n,5], row index; sorted by (isp, id)."""
    ids, sp, rec, row = [], [], [], []
    nsp, nyl = np2.shape
    for isp in range(nsp):
        for jl in range(nyl):
            n = int(np2[isp, jl])
            blk = up[isp, jl, :n, :]
            ids.append(blk[:, 5].copy().view(np.int64))
            sp.append(np.full(n, isp, np.int64))
            rec.append(blk[:, :5].copy())
            row.append(np.full(n, jl, np.int64))
    ids, sp, rec, row = map(np.concatenate, (ids, sp, rec, row))
    order = np.lexsort((ids, sp))
    return ids[order], sp[order], rec[order], row[order]


def oracle_state(w, rank=0):
    """Copies of one rank's host arrays."""
    return dict(up=w.array(rank, O.UP).copy(), gp=w.array(rank, O.GP).copy(), uf=w.array(rank, O.UF).copy(),
                np2=w.array(rank, O.NP2).copy(), cumcnt=w.array(rank, O.CUMCNT).copy(),
                uj=w.array(rank, O.UJ).copy(), df=w.array(rank, O.DF).copy())


def make_world(nx, ny, ppc, nranks=1, steps=0, seed=20260117, **kw):
    prm = O.weibel_params(nx, ny, ppc, nranks=nranks, **kw)
    w = O.World(prm)
    w.ic_weibel(seed)
    if steps:
        w.step(steps)
    return prm, w


def rel_to_max(a, b):
    """max |a-b| per last-axis component relative to that component's max |b| (norm-wise)."""
    d = np.abs(a - b).reshape(-1, a.shape[-1]).max(axis=0)
    s = np.abs(b).reshape(-1, b.shape[-1]).max(axis=0)
    return d / np.where(s > 0, s, 1.0)


def particle_err(rec_a, rec_b, xscale, vth):
    """position error relative to the domain scale, momentum error relative to max(|u|, vth)."""
    ex = np.abs(rec_a[:, :2] - rec_b[:, :2]).max() / xscale
    eu = (np.abs(rec_a[:, 2:] - rec_b[:, 2:]) / np.maximum(np.abs(rec_b[:, 2:]), vth)).max()
    return ex, eu
