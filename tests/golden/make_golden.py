"""Generates the committed golden fixtures from the CPU oracle (parity build).

    python tests/golden/make_golden.py

The reference (Fortran + MPI) cannot be run in this environment and has no fixtures of its own
for the hot path, so these vectors freeze the *oracle's* behaviour: the CPU test reproduces them
bit for bit, the GPU test matches them to the north-star tolerances.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
from helpers import make_world  # noqa: E402


def one(name, nx, ny, ppc, steps, **kw):
    prm, w0 = make_world(nx, ny, ppc, **kw)
    ic = dict(up0=w0.array(0, O.UP).copy(), np20=w0.array(0, O.NP2).copy(), cumcnt0=w0.array(0, O.CUMCNT).copy(),
              uf0=w0.array(0, O.UF).copy())
    w0.step(steps)
    ids, sp, rec = w0.particles_by_id()
    np.savez_compressed(os.path.join(HERE, name), nx=nx, ny=ny, ppc=ppc, steps=steps, ids=ids, sp=sp, rec=rec,
                        uf=w0.array(0, O.UF).copy(), uj=w0.array(0, O.UJ).copy(), cumcnt=w0.array(0, O.CUMCNT).copy(),
                        np2=w0.array(0, O.NP2).copy(), cg_iters=np.array(w0.cg_iters()), energy=w0.energy(), **ic)
    w0.close()


if __name__ == "__main__":
    one("weibel_16x8_p4_s5.npz", 16, 8, 4, 5)
