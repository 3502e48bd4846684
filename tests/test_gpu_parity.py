"""Parity of the CUDA hot path (through the C ABI) against the CPU oracle.  -m gpu.

Tolerances (BASELINE.json north_star): per-cell counts bit-exact; x, y, u matched by
particle ID, J and E/B <= 1e-12 relative after one step (J/E/B relative to each component's
max-abs, conditional on equal CG iteration counts); with WM_FLAG_EXACT_PUSH the push is
bit-identical to the oracle.
"""
import numpy as np
import pytest

import oracle_lib as O
from helpers import (em_wave_in_plasma_setup, flatten_by_id, langmuir_fit, make_langmuir_world, make_world, oracle_state, particle_err,
                     rel_to_max)

pytestmark = pytest.mark.gpu
TOL = 1e-12


def ctx_for(prm, **kw):
    import wumingpic2d_b200 as wm
    return wm.Context.from_params(prm, **kw)


@pytest.fixture(scope="module")
def warm():
    """A 48x24 Weibel world advanced 6 steps: non-trivial fields, particles off their IC lattice.
    nx is not a multiple of the 16-cell tile and ny not of 8: partial tiles are exercised."""
    prm, w = make_world(48 + 8, 24 + 4, 12, steps=6)
    return prm, w


def test_upload_sort_roundtrip(warm):
    """wm_upload_particles == sort__bucket (common/sort.f90:36): cumcnt bit-exact, rows hold the
    same records bit for bit."""
    prm, w = warm
    s = oracle_state(w)
    c = ctx_for(prm)
    rng = np.random.default_rng(1)
    up = s["up"].copy()
    for isp in range(up.shape[0]):  # shuffle every row: arbitrary order in, sorted out
        for jl in range(up.shape[1]):
            n = s["np2"][isp, jl]
            up[isp, jl, :n] = up[isp, jl, rng.permutation(n)]
    c.upload_particles(up, s["np2"])
    out, np2, cum = c.download_particles()
    assert np.array_equal(np2, s["np2"])
    assert np.array_equal(cum, s["cumcnt"])
    a, b = flatten_by_id(out, np2), flatten_by_id(s["up"], s["np2"])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    # cell invariant: int(x) is the cell the slot belongs to
    for isp in range(out.shape[0]):
        for jl in range(out.shape[1]):
            xs = out[isp, jl, :np2[isp, jl], 0].astype(np.int64) - prm["nxgs"]
            assert np.array_equal(np.bincount(xs, minlength=prm["nx"]), np.diff(cum[isp, jl]))
            assert np.all(np.diff(xs) >= 0)
    c.close()


def test_host_sort_bucket_matches_oracle(warm):
    prm, _ = warm
    _, w = make_world(prm["nx"], prm["ny"], prm["n0"], steps=3)
    w.particle_solv(); w.field_fdtd_i(); w.bc_particle_x(); assert w.bc_particle_y() == 0
    s = oracle_state(w)           # gp is now an unsorted set of row lists
    w.sort_bucket()
    ref = oracle_state(w)
    c = ctx_for(prm)
    out = np.zeros_like(s["gp"]); cum = np.zeros_like(s["cumcnt"])
    c.host_sort__bucket(out, s["gp"], cum, s["np2"])
    assert np.array_equal(cum, ref["cumcnt"])
    a, b = flatten_by_id(out, s["np2"]), flatten_by_id(ref["up"], ref["np2"])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    c.close()


@pytest.mark.parametrize("exact", [True, False])
def test_push(warm, exact):
    """particle__solv (common/particle.f90:48): same slot order, so compare slot by slot."""
    import wumingpic2d_b200 as wm
    prm, w0 = warm
    s = oracle_state(w0)
    _, w = make_world(prm["nx"], prm["ny"], prm["n0"], steps=6)
    w.particle_solv()
    ref = w.array(0, O.GP)
    c = ctx_for(prm, flags=wm.WM_FLAG_EXACT_PUSH if exact else 0)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(s["uf"])
    c.particle__solv()
    gp = c.download_gp()
    for isp in range(gp.shape[0]):
        for jl in range(gp.shape[1]):
            n = s["np2"][isp, jl]
            a, b = gp[isp, jl, :n], ref[isp, jl, :n]
            assert np.array_equal(a[:, 5].view(np.int64), b[:, 5].view(np.int64))
            if exact:
                assert np.array_equal(a.view(np.int64), b.view(np.int64)), "EXACT push must be bit-identical to the CPU path"
            else:
                ex, eu = particle_err(a[:, :5], b[:, :5], prm["nx"], prm["vte"])
                assert ex <= TOL and eu <= TOL
    c.close()


def test_deposit_and_field_solve(warm):
    """ele_cur (field.f90:189), bc__curre, then the whole field__fdtd_i (field.f90:66)."""
    import wumingpic2d_b200 as wm
    prm, w0 = warm
    s = oracle_state(w0)
    _, w = make_world(prm["nx"], prm["ny"], prm["n0"], steps=6)
    w.particle_solv()
    w.ele_cur()
    uj_raw = w.array(0, O.UJ).copy()
    w.bc_curre()
    uj_bc = w.array(0, O.UJ).copy()
    c = ctx_for(prm, flags=wm.WM_FLAG_EXACT_PUSH)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(s["uf"])
    # bring the CG warm start (the reference's SAVEd df) to the same state: run the same history
    c2 = None
    c.particle__solv()
    c.field__ele_cur()
    assert rel_to_max(c.download_current(), uj_raw).max() <= TOL
    c.bc__curre()
    assert rel_to_max(c.download_current(), uj_bc).max() <= TOL
    c.close()


def test_stage_calls_one_step(warm):
    """The five calls of proj/weibel/app.f90:102-107 one by one, from the same history (so the CG
    warm start df matches), against the oracle."""
    import wumingpic2d_b200 as wm
    prm, _ = warm
    _, w = make_world(prm["nx"], prm["ny"], prm["n0"])
    s = oracle_state(w)
    c = ctx_for(prm, flags=wm.WM_FLAG_EXACT_PUSH)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(s["uf"])
    for it in range(4):
        w.step(1)
        c.particle__solv(); c.field__fdtd_i(); c.bc__particle_x(); c.bc__particle_y(); c.sort__bucket()
        assert c.cg_iters() == w.cg_iters()
        uf = c.download_field()
        assert rel_to_max(uf, w.array(0, O.UF)).max() <= TOL
        up, np2, cum = c.download_particles()
        assert np.array_equal(cum, w.array(0, O.CUMCNT)), "per-cell counts must be bit-exact"
        a, b = flatten_by_id(up, np2), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
        ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
        assert ex <= TOL and eu <= TOL
    c.close()


@pytest.mark.parametrize("exact", [True, False])
def test_fused_step_matches_oracle(exact):
    """wm_step (push+deposit+boundary fused, field solve, scatter) against the oracle, config-1-like
    physics (proj/weibel/config_sample.json scaled down)."""
    import wumingpic2d_b200 as wm
    prm, w = make_world(64, 32, 20)
    s = oracle_state(w)
    c = ctx_for(prm, flags=wm.WM_FLAG_EXACT_PUSH if exact else 0)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(s["uf"])
    w.step(1)
    c.step(1)
    assert c.cg_iters() == w.cg_iters()
    assert rel_to_max(c.download_field(), w.array(0, O.UF)).max() <= TOL
    assert rel_to_max(c.download_current(), w.array(0, O.UJ)).max() <= TOL
    up, np2, cum = c.download_particles()
    assert np.array_equal(cum, w.array(0, O.CUMCNT)), "per-cell counts must be bit-exact"
    a, b = flatten_by_id(up, np2), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
    assert np.array_equal(a[0], b[0])
    if exact:
        assert np.array_equal(a[2].view(np.int64), b[2].view(np.int64)), "first step from identical fields: bit-identical particles"
    ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
    assert ex <= TOL and eu <= TOL
    # a few more steps: still within tolerance (errors grow slowly before chaos sets in)
    w.step(4)
    c.step(4)
    up, np2, cum = c.download_particles()
    a, b = flatten_by_id(up, np2), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
    ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
    assert ex <= 1e-10 and eu <= 1e-10
    assert rel_to_max(c.download_field(), w.array(0, O.UF)).max() <= 1e-10
    c.close()


def test_gauss_law_and_energy_history():
    """Discrete charge conservation (div E - 4 pi rho constant to roundoff) and the energy history
    of proj/weibel/app.f90:479-545 over 100 steps against the oracle."""
    prm, w = make_world(64, 64, 10)
    s = oracle_state(w)
    c = ctx_for(prm)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(s["uf"])
    e0 = c.energy()
    assert np.allclose(e0, w.energy(), rtol=1e-13, atol=0)
    for _ in range(4):
        w.step(25)
        c.step(25)
        eg, eo = c.energy(), w.energy()
        assert abs(eg.sum() - eo.sum()) <= 1e-9 * eo.sum()
        # load the device state into a scratch oracle world to evaluate the Gauss residual
        up, np2, cum = c.download_particles()
        g = O.World(prm)
        g.array(0, O.UP)[...] = up
        g.array(0, O.NP2)[...] = np2
        g.array(0, O.UF)[...] = c.download_field()
        res, scale = g.gauss_residual()
        assert res <= 1e-12 * scale
        g.close()
        # the device-side diagnostic evaluates the same definition
        rd, sd = c.gauss_residual()
        assert abs(sd - scale) <= 1e-12 * scale and rd <= 1e-12 * sd
    c.close()


def test_host_step_dropin():
    """wm_host_step: one step on host arrays in the reference's layout (the e2e path)."""
    prm, w = make_world(32, 16, 8)
    s = oracle_state(w)
    c = ctx_for(prm)
    up, uf, np2, cum = s["up"].copy(), s["uf"].copy(), s["np2"].copy(), s["cumcnt"].copy()
    for _ in range(3):
        w.step(1)
        c.host_step(up, uf, np2, cum)
    assert np.array_equal(cum, w.array(0, O.CUMCNT))
    assert rel_to_max(uf, w.array(0, O.UF)).max() <= 1e-11
    a, b = flatten_by_id(up, np2), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
    assert np.array_equal(a[0], b[0])
    ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
    assert ex <= 1e-11 and eu <= 1e-11
    c.close()


def test_host_steps_sync_interval():
    """wm_host_steps(n): one upload, n steps, one download -- what the shim does with WM_SYNC_INTERVAL = n (host arrays valid
    every n-th step).  Here the CG warm start survives between the steps, as in a resident run, so it equals n oracle steps."""
    prm, w = make_world(32, 16, 8)
    s = oracle_state(w)
    c = ctx_for(prm)
    up, uf, np2, cum = s["up"].copy(), s["uf"].copy(), s["np2"].copy(), s["cumcnt"].copy()
    w.step(5)
    c.host_steps(up, uf, np2, cum, 5)
    assert c.cg_iters() == w.cg_iters()
    assert np.array_equal(cum, w.array(0, O.CUMCNT)) and np.array_equal(np2, w.array(0, O.NP2))
    assert rel_to_max(uf, w.array(0, O.UF)).max() <= 1e-10
    a, b = flatten_by_id(up, np2), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
    assert np.array_equal(a[0], b[0])
    ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
    assert ex <= 1e-10 and eu <= 1e-10
    c.close()


def test_moments_match_oracle(warm):
    prm, w0 = warm
    s = oracle_state(w0)
    _, w = make_world(prm["nx"], prm["ny"], prm["n0"], steps=6)
    w.mom_accl(); w.mom_nvt(); w.bc_mom()
    c = ctx_for(prm)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(s["uf"])
    mom = c.moments()
    ref = w.array(0, O.MOM)
    # interior nodes only: ghosts hold pre-fold partial sums in both implementations
    assert rel_to_max(mom[:, 1:-1, 1:-1], ref[:, 1:-1, 1:-1]).max() <= 1e-12
    # the same through the three reference procedures one by one (proj/weibel/app.f90:121-123)
    _, w2 = make_world(prm["nx"], prm["ny"], prm["n0"], steps=6)
    w2.mom_accl(); w2.mom_nvt()
    c.mom_calc__accl()
    raw = c.mom_calc__nvt()
    assert rel_to_max(raw, w2.array(0, O.MOM)).max() <= 1e-12     # ghosts included, before the fold
    c.bc__mom(raw)
    assert rel_to_max(raw[:, 1:-1, 1:-1], ref[:, 1:-1, 1:-1]).max() <= 1e-12
    c.close()


def test_device_ic_matches_oracle_ic():
    """wm_ic_weibel uses the oracle's stream definition: positions bit-identical, velocities to a few
    ulp (device libm), analytic cumcnt identical."""
    prm, w = make_world(32, 16, 8)
    c = ctx_for(prm)
    c.ic_weibel(20260117, prm["n0"], prm["vti"], prm["vte"], prm["t_ani"], prm["b0"])
    up, np2, cum = c.download_particles()
    assert np.array_equal(cum, w.array(0, O.CUMCNT)) and np.array_equal(np2, w.array(0, O.NP2))
    a, b = flatten_by_id(up, np2), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
    assert np.array_equal(a[0], b[0])
    assert np.array_equal(a[2][:, :2], b[2][:, :2])
    assert np.abs(a[2][:, 2:] - b[2][:, 2:]).max() <= 1e-14
    c.close()


def test_call_order_is_enforced(warm):
    import wumingpic2d_b200 as wm
    prm, w = warm
    c = ctx_for(prm)
    with pytest.raises(wm.WmError):
        c.particle__solv()          # nothing uploaded
    s = oracle_state(w)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    with pytest.raises(wm.WmError):
        c.sort__bucket()            # boundary not applied yet
    c.close()


@pytest.mark.parametrize("env", [{"WM_INPLACE": "0"}, {"WM_SLACK": "0.4"}, {"WM_SLACK": "12"}, {"WM_SM": "0"},
                                 {"WM_CG3": "1", "WM_CG": "0"}, {"WM_CG": "0"}, {"WM_OVERLAP": "0"}, {"WM_SM": "3"}, {"WM_RIMPLACE": "0"},
                                 {"WM_SM": "3", "WM_SLACK": "0.4"}, {"WM_SM": "5"}, {"WM_SM": "5", "WM_SLACK": "0.4"},
                                 {"WM_SM": "5", "WM_OVERLAP": "0"}, {"WM_SLACK": "0.4", "WM_OVERLAP": "0"}])
def test_sort_variants_match_oracle(env, monkeypatch):
    """wm_step with (a) the tag + scatter sort, (b) the in-place sort with so little segment slack that
    segments overflow and the layout is rebuilt nearly every step, (c) generous slack, (d) k_fused<INPLACE> (65
    register sums per lane) instead of k_fused_sm, (e) the host-loop CG with the three-kernel / two-kernel iteration instead of
    the persistent cooperative kernel, (f) everything on one
    stream, (g) k_fused_sm without its in-tile tail (every cell changer through k_place + k_mark_dead), (h) the tail
    with the general k_place instead of k_place_rim, (i) = (g) with overflowing segments: per-cell counts bit-exact and
    particles/fields within tolerance in every case; (j) k_fused_dp (WM_SM=5: ping-pong stores, cell changers placed directly,
    k_place_rim2 for the window rim, k_normalize before every download) with and without overflows.  (The default path --
    k_fused_sm<TAIL> + k_place_rim -- is what every other test runs; (b) drives its overflow branch.)"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    prm, w = make_world(40, 24, 16)
    s = oracle_state(w)
    c = ctx_for(prm)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(s["uf"])
    for _ in range(3):
        w.step(4)
        c.step(4)
        up, np2, cum = c.download_particles()
        assert np.array_equal(cum, w.array(0, O.CUMCNT)), "per-cell counts must be bit-exact"
        a, b = flatten_by_id(up, np2), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
        ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
        assert ex <= 1e-9 and eu <= 1e-9
        assert rel_to_max(c.download_field(), w.array(0, O.UF)).max() <= 1e-9
    if env.get("WM_SLACK") == "0.4":  # noqa
        assert c.rebuilds() > 0, "the overflow -> rebuild path was not exercised"
    c.close()


@pytest.mark.parametrize("kind", ["periodic", "reconnection", "shock"])
def test_persistent_cg_many_blocks(kind):
    """cgm (common/field.f90:319-461) as the persistent cooperative kernel on a grid that is split over many CTAs in both
    directions (block edges, the halo ring through the global r array, wrapped / wall columns): equal CG iteration counts
    per component and fields <= 1e-12 against the oracle step by step, and the same against the host-loop CG."""
    from helpers import make_shock_world, make_wall_world
    import wumingpic2d_b200 as wm
    if kind == "periodic":
        prm, w = make_world(208, 150, 3)
    elif kind == "reconnection":
        prm, w = make_wall_world(160, 128, 3)
    else:
        prm, w = make_shock_world(160, 128, 3, u0=-0.3)
    s = oracle_state(w)
    c = ctx_for(prm)
    assert c.cg_path() == 1, "the persistent CG kernel must be the default on one rank"
    plan = wm.api.cg_plan(prm["nx"], prm["ny"])
    assert plan[0] > 1 and plan[1] > 1, plan
    if kind == "shock":
        c.set_u_inject(-0.3)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(s["uf"])
    for it in range(4):
        w.step(1)
        c.step(1)
        assert c.cg_iters() == w.cg_iters(), (it, c.cg_iters(), w.cg_iters())
        assert max(c.cg_iters()) > 0
        assert rel_to_max(c.download_field(), w.array(0, O.UF)).max() <= (TOL if it == 0 else 1e-10)
        assert rel_to_max(c.download_dfield()[..., :3], w.array(0, O.DF)[..., :3]).max() <= (1e-11 if it == 0 else 1e-9)
    up, np2, cum = c.download_particles()
    assert np.array_equal(cum, w.array(0, O.CUMCNT))
    c.close()
    w.close()


def test_benchmark_strip_matches_oracle():
    """The benchmark's own initial condition and row length against the oracle: a 4096 x 8-row strip of BASELINE configs[1]
    (64 ppc x 2 species, 4.2 M particles: 256 tiles wide, tiles at both slab edges, the mover-queue drain path at 64 ppc).
    The device generates the IC itself (wm_ic_weibel, what bench.py times) and must agree with the oracle's generator to
    a few ulp; then both start from the oracle's state: counts bit-exact, particles by ID / fields <= 1e-12 after one
    step, <= 1e-10 after three."""
    import wumingpic2d_b200 as wm
    prm = O.weibel_params(4096, 8, 64, cap_factor=1.25)
    w = O.World(prm, fast=False)
    w.ic_weibel(20260117)
    s = oracle_state(w)
    c = ctx_for(prm)
    c.ic_weibel(20260117, 64, prm["vti"], prm["vte"], prm["t_ani"], prm["b0"])
    up, np2, cum = c.download_particles()
    assert np.array_equal(cum, s["cumcnt"]) and np.array_equal(np2, s["np2"])
    a, b = flatten_by_id(up, np2), flatten_by_id(s["up"], s["np2"])
    assert np.array_equal(a[0], b[0])
    ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
    assert ex <= 1e-15 and eu <= 1e-12, (ex, eu)       # positions identical, Box-Muller in device libm: a few ulp
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(s["uf"])
    for it in range(3):
        w.step(1)
        c.step(1)
        assert c.cg_iters() == w.cg_iters()
        up, np2, cum = c.download_particles()
        assert np.array_equal(cum, w.array(0, O.CUMCNT)), "per-cell counts must be bit-exact (step %d)" % it
        a, b = flatten_by_id(up, np2), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
        assert np.array_equal(a[0], b[0])
        ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
        tol = TOL if it == 0 else 1e-10
        assert ex <= tol and eu <= tol, (it, ex, eu)
        assert rel_to_max(c.download_field(), w.array(0, O.UF)).max() <= tol
        assert rel_to_max(c.download_current(), w.array(0, O.UJ)).max() <= tol
    assert c.rebuilds() == 0
    c.close(); w.close()


def test_langmuir_oscillation_on_device():
    """A known answer for the CUDA path itself, not through the oracle: cold uniform plasma, heavy ions, electrons on a
    quiet lattice with ux = v0 sin(k x) (the oracle is only the container of the initial condition).  260 wm_step on the
    device must make Ex oscillate at the plasma frequency the set-up asked for (0.994 omega_pe with the spline shape
    factors; tolerance 3 %), with the cold-plasma amplitude e E_max = m v0 omega_pe (2 %) and no growth or damping
    (tests/test_oracle_pins.py::test_known_answer_langmuir_frequency is the same check on the oracle)."""
    nx, wpe, v0 = 32, 0.1, 1e-3
    prm, w = make_langmuir_world(nx, 4, 16, v0=v0, m=1, wpe=wpe)
    s = oracle_state(w)
    w.close()
    c = ctx_for(prm)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(s["uf"])
    series = []
    for n in range(260):
        c.step(1)
        series.append(c.download_field()[2:-2, 2 + nx // 4, 3].mean())
    c.close()
    omega, peaks, ncross = langmuir_fit(series, prm["delt"])
    assert ncross >= 6
    assert abs(omega / wpe - 1.0) <= 0.03, omega
    assert 0.8 <= peaks[-1] / peaks[0] <= 1.2, peaks
    e_max = prm["r"][1] * v0 * wpe / abs(prm["q"][1])
    assert abs(peaks[0] / e_max - 1.0) <= 0.02, (peaks[0], e_max)


@pytest.mark.parametrize("m,cfl,gfac", [(1, 1.0, 0.501), (3, 0.5, 0.75)])
def test_vacuum_wave_amplification_on_device(m, cfl, gfac):
    """The second known answer for the CUDA path itself: a standing wave Ez = cos(k x), B = 0 in (effectively) vacuum --
    the particles carry a charge of 1e-14 -- must be multiplied per wm_step by Re(G^n), G = (1 + i s (1 - gfac)) /
    (1 - i s gfac), s = (c dt / delx) 2 sin(k delx / 2): the amplification factor of the implicit scheme of
    field.f90:125-171 (derivation in tests/test_oracle_pins.py::test_known_answer_vacuum_wave_dispersion).  Checks
    k_rhs, the CG kernels, k_efield, the halo fills and k_update_uf against the algebra of the scheme."""
    nx, ny, nsteps = 24, 8, 24
    prm = O.weibel_params(nx, ny, 2, cfl=cfl, gfac=gfac)
    prm["q"] = [1e-14, -1e-14]
    w = O.World(prm)
    w.ic_weibel(3)
    s = oracle_state(w)
    w.close()
    k = 2 * np.pi * m / nx
    uf = np.zeros_like(s["uf"])
    ii = np.arange(uf.shape[1]) + (prm["nxgs"] - 2)
    ez0 = np.cos(k * (ii + 0.5))[None, :] * np.ones((uf.shape[0], 1))
    uf[:, :, 5] = ez0
    c = ctx_for(prm)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(uf)
    s_ = prm["c"] * prm["delt"] / prm["delx"] * 2 * np.sin(k * prm["delx"] / 2)
    G = (1 + 1j * s_ * (1 - gfac)) / (1 - 1j * s_ * gfac)
    Gn = 1.0 + 0j
    for n in range(1, nsteps + 1):
        c.step(1)
        Gn *= G
        f = c.download_field()[2:-2, 2:-2]
        # CG tolerance 1e-6 per solve (field.f90:339), accumulated over the steps
        assert np.abs(f[:, :, 5] - Gn.real * ez0[2:-2, 2:-2]).max() <= 2e-5, n
        assert abs(np.abs(f[:, :, 1]).max() - abs(Gn.imag)) <= 2e-5, n
        for comp in (0, 2, 3, 4):
            assert np.abs(f[:, :, comp]).max() <= 1e-10, (n, comp)
    c.close()


def test_light_wave_in_plasma_on_device():
    """Third known answer for the CUDA path itself: an Ez standing wave in a cold plasma at rest oscillates at
    omega^2 = omega_pe^2 + c^2 k^2 (2 %), and at the scheme's own dispersion relation to 0.3 % -- the uz push, the Jz
    deposit and the implicit solve together (tests/test_oracle_pins.py::test_known_answer_light_wave_in_plasma)."""
    prm, w, uf, om_scheme, om_text = em_wave_in_plasma_setup(2)
    s = oracle_state(w)
    w.close()
    c = ctx_for(prm)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(uf)
    series = []
    for n in range(300):
        c.step(1)
        series.append(c.download_field()[2:-2, 2, 5].mean())
    c.close()
    omega, peaks, ncross = langmuir_fit(series, prm["delt"])
    assert ncross >= 15
    assert abs(omega / om_scheme - 1.0) <= 3e-3, (omega, om_scheme)
    assert abs(omega / om_text - 1.0) <= 0.02, (omega, om_text)
    assert 0.85 <= peaks[-1] / peaks[0] <= 1.05, peaks


def test_dense_cells_drain_early():
    """k_fused_sm queues the movers of a cell (40 slots per cell) and drains the queue when the cell is done.  With
    200 particles per cell per species about 60 particles per cell change cell in a step, so the queue is drained
    early and the particle loop re-entered several times per cell: J (hence E, B) and the particles must still
    match the oracle, and the discrete Gauss law must hold."""
    prm, w = make_world(24, 16, 200)
    s = oracle_state(w)
    c = ctx_for(prm)
    c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    c.upload_field(s["uf"])
    for it in range(3):
        w.step(1)
        c.step(1)
        assert c.cg_iters() == w.cg_iters()
        tol = TOL if it == 0 else 1e-10
        assert rel_to_max(c.download_current(), w.array(0, O.UJ)).max() <= tol
        assert rel_to_max(c.download_field(), w.array(0, O.UF)).max() <= tol
        up, np2, cum = c.download_particles()
        assert np.array_equal(cum, w.array(0, O.CUMCNT)), "per-cell counts must be bit-exact"
        a, b = flatten_by_id(up, np2), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
        assert np.array_equal(a[0], b[0])
        ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
        assert ex <= tol and eu <= tol
    c.close()


@pytest.mark.parametrize("env", [{}, {"WM_INPLACE": "0"}])
def test_reconnection_walls_match_oracle(env, monkeypatch):
    """WM_BC_RECONNECTION (proj/reconnection/boundary_reconnection.f90): reflecting particle walls at
    nxs+1 / nxe-1 with momentum flip, conducting-wall rules for df and the CG vectors, no x fold of the
    current; fused step and the stage calls against the oracle's restatement of that module."""
    import wumingpic2d_b200 as wm
    from helpers import make_wall_world
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    prm, w = make_wall_world(40, 20, 10)
    s = oracle_state(w)
    for flags in (0, wm.WM_FLAG_EXACT_PUSH):
        _, w = make_wall_world(40, 20, 10)
        c = ctx_for(prm, flags=flags)
        c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
        c.upload_field(s["uf"])
        for it in range(8):
            w.step(1)
            if it == 5:   # one step through the five stage calls (bc__particle_x = k_bcx reflects)
                c.particle__solv(); c.field__fdtd_i(); c.bc__particle_x(); c.bc__particle_y(); c.sort__bucket()
            else:
                c.step(1)
            assert c.cg_iters() == w.cg_iters()
            up, np2, cum = c.download_particles()
            assert np.array_equal(cum, w.array(0, O.CUMCNT)), "per-cell counts must be bit-exact (step %d)" % it
            a, b = flatten_by_id(up, np2), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
            ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
            tol = 1e-12 if it == 0 else 1e-10
            assert ex <= tol and eu <= tol
            assert rel_to_max(c.download_field(), w.array(0, O.UF)).max() <= tol
        w.mom_accl(); w.mom_nvt(); w.bc_mom()
        mom = c.moments()
        assert rel_to_max(mom[:, 1:-1, 1:-1], w.array(0, O.MOM)[:, 1:-1, 1:-1]).max() <= 1e-10
        c.close(); w.close()


def test_shock_injection_matches_oracle():
    """WM_BC_SHOCK (proj/shock/boundary_shock.f90): reflecting left wall, injection wall at xend with ux -> 2 u0 - ux,
    applied BEFORE the deposit (proj/shock/app.f90:112-113), df = 0 in the right ghost column; fused step (FMA and
    exact push) and the stage calls against the oracle's restatement."""
    import wumingpic2d_b200 as wm
    from helpers import make_shock_world
    u0 = -0.3
    prm, w0 = make_shock_world(40, 20, 10, u0=u0)
    s = oracle_state(w0)
    for flags in (0, wm.WM_FLAG_EXACT_PUSH):
        _, w = make_shock_world(40, 20, 10, u0=u0)
        c = ctx_for(prm, flags=flags, bc=wm.WM_BC_SHOCK)
        c.set_u_inject(u0)
        c.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
        c.upload_field(s["uf"])
        nref = 0
        for it in range(8):
            w.step(1)
            if it == 5:   # one step through the stage calls of proj/shock/app.f90:111-116
                c.particle__solv(); c.bc__injection(u0); c.field__fdtd_i(); c.bc__particle_y(); c.sort__bucket()
            else:
                c.step(1)
            assert c.cg_iters() == w.cg_iters()
            up, np2, cum = c.download_particles()
            assert np.array_equal(cum, w.array(0, O.CUMCNT)), "per-cell counts must be bit-exact (step %d)" % it
            a, b = flatten_by_id(up, np2), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[3], b[3])
            ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
            tol = 1e-12 if it == 0 else 1e-10
            assert ex <= tol and eu <= tol
            assert rel_to_max(c.download_field(), w.array(0, O.UF)).max() <= tol
            assert rel_to_max(c.download_current(), w.array(0, O.UJ)).max() <= tol
        # the walls were actually hit: some particles have ux > 0 although the plasma drifts to the left at -0.3
        assert (b[2][:, 2] > 0.2).sum() > 0
        c.close(); w.close()


def test_shock_moving_box_matches_oracle():
    """The shock app works on the active range nxs..nxe and moves nxe (`relocate`, proj/shock/app.f90:611-680): the nxs,
    nxe arguments of particle__solv, bc__injection, field__fdtd_i, sort__bucket vary in time.  Start with a short box,
    run, grow the box twice the way relocate does (new particles in the cell nxe-1, upstream fields in the new
    columns; applied to the oracle's arrays and to downloaded device arrays alike), and compare after every step."""
    import wumingpic2d_b200 as wm
    from helpers import make_shock_world, shock_relocate
    u0, b0, n0 = -0.3, 0.02, 10
    nx, ny = 40, 16
    prm, w = make_shock_world(nx, ny, n0, u0=u0, b0=b0, nxe=2 + 30)
    nxgs = prm["nxgs"]
    nxe = nxgs + 30
    s = oracle_state(w)
    c = ctx_for(prm, bc=wm.WM_BC_SHOCK)
    c.set_xrange(nxgs, nxe)
    c.set_u_inject(u0)
    c.upload_particles(s["up"], s["np2"])
    c.upload_field(s["uf"])

    def compare(tol):
        na = nxe - nxgs + 1
        assert c.cg_iters() == w.cg_iters()
        up, np2, cum = c.download_particles()
        assert np.array_equal(np2, w.array(0, O.NP2))
        assert np.array_equal(cum[:, :, :na + 1], w.array(0, O.CUMCNT)[:, :, :na + 1]), "per-cell counts must be bit-exact"
        a, b = flatten_by_id(up, np2), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
        assert np.array_equal(a[0], b[0])
        ex, eu = particle_err(a[2], b[2], prm["nx"], prm["vte"])
        assert ex <= tol and eu <= tol
        assert rel_to_max(c.download_field(), w.array(0, O.UF)).max() <= tol
        assert rel_to_max(c.download_current(), w.array(0, O.UJ)).max() <= tol

    for it in range(3):
        w.step(1); c.step(1)
        compare(1e-12 if it == 0 else 1e-10)
    for grow in range(2):
        nxe += 1
        # oracle side: edit its arrays in place, as the driver does
        shock_relocate(prm, w.array(0, O.UP), w.array(0, O.NP2), w.array(0, O.UF), nxe, n0, u0, b0, seed=5)
        cum = w.array(0, O.CUMCNT)
        cum[:, :, nxe - nxgs] = cum[:, :, nxe - 1 - nxgs] + n0      # cumcnt(nxe,j,isp) = cumcnt(nxe-1,j,isp) + n0
        assert w.lib.orc_set_xrange(w.h, nxgs, nxe) == 0
        # device side: first growth through downloaded arrays + upload, second through wm_append_particles
        up, np2, _ = c.download_particles()
        uf = c.download_field()
        added = shock_relocate(prm, up, np2, uf, nxe, n0, u0, b0, seed=5)
        c.set_xrange(nxgs, nxe)
        if grow == 0:
            c.upload_particles(up, np2)
        else:
            for isp in range(2):
                c.append_particles(isp, added[isp])
        c.upload_field(uf)
        for it in range(3):
            w.step(1); c.step(1)
            compare(1e-10)
    c.close(); w.close()
