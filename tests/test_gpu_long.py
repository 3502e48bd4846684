"""Long-run parity (north_star: "stay within a stated tolerance on energy history over 1000
steps") and the committed golden fixture.  -m gpu.

Stated tolerance: trajectories decorrelate chaotically, so at long times only
integrated quantities are compared: total energy GPU vs oracle <= 1e-6 relative at every
sampled step up to step 1000, each energy component (kinetic per species, E^2/8pi, B^2/8pi)
<= 1e-3 relative to the total, and the GPU's total-energy drift within the oracle's own
drift +-10 % (+ 1e-9 absolute slack).
"""
import os

import numpy as np
import pytest

import oracle_lib as O
from helpers import flatten_by_id, make_world, oracle_state, particle_err, rel_to_max

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_energy_history_1000_steps():
    import wumingpic2d_b200 as wm
    # proj/weibel/config_sample.json physics (t_ani = 5, v_th = 0.1, omega_pe = 0.1), 64 x 64, 20 ppc
    prm, w = make_world(64, 64, 20)
    s = oracle_state(w)
    c = wm.Context.from_params(prm)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(s["uf"])
    e0 = w.energy().sum()
    assert abs(c.energy().sum() - e0) <= 1e-13 * e0
    for k in range(20):                      # energy every 50 steps, like intvl_mom = 50
        w.step(50)
        c.step(50)
        eo, eg = w.energy(), c.energy()
        assert abs(eg.sum() - eo.sum()) <= 1e-6 * eo.sum(), "step %d" % (50 * (k + 1))
        assert np.abs(eg - eo).max() <= 1e-3 * eo.sum()
    drift_o, drift_g = eo.sum() - e0, eg.sum() - e0
    assert abs(drift_g - drift_o) <= 0.1 * abs(drift_o) + 1e-9 * e0
    # the Weibel instability has grown: magnetic energy is no longer zero
    assert eg[-1] > 1e-4 * e0
    # particle number and per-cell count bookkeeping survived 1000 sorts
    up, np2, cum = c.download_particles(want_up=False)
    assert int(np2.sum()) == 2 * 64 * 64 * 20
    c.close(); w.close()


def test_golden_fixture_on_gpu():
    """The committed fixture (tests/golden/make_golden.py): 5 steps from the stored initial state."""
    import wumingpic2d_b200 as wm
    g = np.load(os.path.join(HERE, "golden", "weibel_16x8_p4_s5.npz"))
    prm = O.weibel_params(int(g["nx"]), int(g["ny"]), int(g["ppc"]))
    for flags, tol in ((wm.WM_FLAG_EXACT_PUSH, 1e-11), (0, 1e-10)):
        c = wm.Context.from_params(prm, flags=flags)
        c.upload_particles_sorted(np.ascontiguousarray(g["up0"]), np.ascontiguousarray(g["np20"]), np.ascontiguousarray(g["cumcnt0"]))
        c.upload_field(np.ascontiguousarray(g["uf0"]))
        c.step(int(g["steps"]))
        assert c.cg_iters() == list(g["cg_iters"])
        up, np2, cum = c.download_particles()
        assert np.array_equal(cum, g["cumcnt"]), "per-cell counts must be bit-exact"
        ids, sp, rec, _ = flatten_by_id(up, np2)
        assert np.array_equal(ids, g["ids"]) and np.array_equal(sp, g["sp"])
        ex, eu = particle_err(rec, g["rec"], prm["nx"], prm["vte"])
        assert ex <= tol and eu <= tol
        assert rel_to_max(c.download_field(), g["uf"]).max() <= tol
        assert np.allclose(c.energy(), g["energy"], rtol=1e-10, atol=0)
        c.close()
