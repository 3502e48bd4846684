"""Small worlds for compute-sanitizer (memcheck / racecheck / synccheck): a few wm_step calls of one boundary
kind on a grid with partial tiles, checked against the oracle's per-cell counts so a silent corruption shows.

    compute-sanitizer --tool racecheck python scripts/sanitize_driver.py weibel 3

Variants are selected with the library's environment switches (WM_SM, WM_SLACK, WM_INPLACE, WM_CG ...).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import oracle_lib as O  # noqa: E402
from helpers import make_shock_world, make_wall_world, make_world  # noqa: E402


def main():
    import wumingpic2d_b200 as wm
    kind = sys.argv[1] if len(sys.argv) > 1 else "weibel"
    nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    if kind == "weibel":
        prm, w = make_world(48 + 8, 24 + 4, 12, steps=2)
    elif kind == "reconnection":
        prm, w = make_wall_world(40, 19, 8)
    elif kind == "shock":
        prm, w = make_shock_world(40, 19, 8, u0=-0.3)
    elif kind == "hostpipe":
        # the pipelined wm_host_step: six chunks of one tile row (the last one ragged), row-range kernels, four streams
        os.environ["WM_HOSTPIPE_ROWS"] = "8"
        prm, w = make_world(48 + 8, 44, 10)
        c = wm.Context.from_params(prm, device=0)
        up, uf = w.array(0, O.UP).copy(), w.array(0, O.UF).copy()
        np2, cum = w.array(0, O.NP2).copy(), w.array(0, O.CUMCNT).copy()
        for _ in range(nsteps):
            w.step(1)
            c.host_step(up, uf, np2, cum)
            assert c.host_pipe_chunks() == 6
            assert np.array_equal(cum, w.array(0, O.CUMCNT)) and np.array_equal(np2, w.array(0, O.NP2)), "per-cell counts differ from the oracle"
        print("sanitize_driver hostpipe: %d steps ok" % nsteps)
        c.close()
        w.close()
        return
    else:
        raise SystemExit("kind: weibel | reconnection | shock | hostpipe")
    c = wm.Context.from_params(prm, device=0)
    if kind == "shock":
        c.set_u_inject(-0.3)
    c.upload_particles_sorted(w.array(0, O.UP).copy(), w.array(0, O.NP2).copy(), w.array(0, O.CUMCNT).copy())
    c.upload_field(w.array(0, O.UF).copy())
    for _ in range(nsteps):
        w.step(1)
        c.step(1)
    up, np2, cum = c.download_particles()
    assert np.array_equal(cum, w.array(0, O.CUMCNT)), "per-cell counts differ from the oracle"
    mom = c.moments()
    e = c.energy()
    assert np.isfinite(mom).all() and np.isfinite(e).all()
    print("sanitize_driver %s: %d steps ok, rebuilds=%d, env=%s" % (
        kind, nsteps, c.rebuilds(), {k: v for k, v in os.environ.items() if k.startswith("WM_")}))
    c.close()
    w.close()


if __name__ == "__main__":
    main()
