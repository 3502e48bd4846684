"""Restart snapshots in the reference's NAME.json + NAME.raw format (common/paraio.f90:93-426)."""
import json
import os
import sys

import numpy as np
import pytest

import oracle_lib as O
from helpers import flatten_by_id, make_world, oracle_state, particle_err, rel_to_max


def _cfg(prm):
    return dict(nxgs=prm["nxgs"], nxge=prm["nxgs"] + prm["nx"] - 1, nygs=prm["nygs"], nyge=prm["nygs"] + prm["ny"] - 1,
                delx=prm["delx"], delt=prm["delt"], c=prm["c"], r=prm["r"], q=prm["q"])


def test_restart_file_layout_and_roundtrip(tmp_path):
    """Two ranks' arrays -> files -> back, bit for bit; byte layout and JSON schema as paraio__output writes them:
    attributes in the reference's order at a running displacement, datasets rank-major with column-major shapes."""
    from wumingpic2d_b200 import snapshot as S
    prm, w = make_world(16, 12, 3, nranks=2, steps=2)
    ups = [w.array(r, O.UP).copy() for r in range(2)]
    np2s = [w.array(r, O.NP2).copy() for r in range(2)]
    ufs = [w.array(r, O.UF).copy() for r in range(2)]
    base = str(tmp_path / "0000002_restart")
    S.write_restart(base, 2, prm["nxgs"], prm["nxgs"] + 15, _cfg(prm), ups, np2s, ufs)
    root = json.load(open(base + ".json"))
    assert list(root) == ["meta", "attribute", "dataset"]
    assert root["meta"]["rawfile"] == "0000002_restart.raw"
    # mpiio_get_endian_flag (utils/iocore/mpiio.f90:97-109): 1 on a little-endian writer, 16777216 on a big-endian one
    assert root["meta"]["endian"] == (1 if sys.byteorder == "little" else 16777216)
    names = list(root["attribute"])
    assert names == ["dummy_attribute", "it", "nxs", "nxe", "ndim", "np", "nxgs", "nxge", "nygs", "nyge", "nsp", "nproc",
                     "delx", "delt", "c", "r", "q"]
    for rec in root["attribute"].values():
        assert set(rec) == {"datatype", "offset", "size", "ndim", "shape", "description", "data"}
    off = 0
    for n in names:   # running displacement, no padding
        assert root["attribute"][n]["offset"] == off
        off += root["attribute"][n]["size"]
    assert root["attribute"]["dummy_attribute"]["data"] == 8 and root["attribute"]["nproc"]["data"] == 2
    nyl, npcap = ups[0].shape[1], ups[0].shape[2]
    d = root["dataset"]
    assert d["up"]["shape"] == [6, npcap, nyl, 2, 2] and d["up"]["offset"] == off and d["up"]["size"] == 2 * ups[0].nbytes
    assert d["np2"]["shape"] == [nyl, 2, 2] and d["np2"]["datatype"] == "i4"
    assert d["uf"]["shape"] == [6, 16 + 4, nyl + 4, 2]
    assert os.path.getsize(base + ".raw") == d["uf"]["offset"] + d["uf"]["size"]
    # the second rank's first particle sits where the Fortran element up(1,1,nys,1) of rank 1 would be
    raw = np.fromfile(base + ".raw", dtype=np.float64, count=6, offset=d["up"]["offset"] + ups[0].nbytes)
    assert np.array_equal(raw.view(np.int64), ups[1][0, 0, 0].view(np.int64))   # (the id is a bit pattern: NaN as a double)
    attrs, up2, np22, uf2 = S.read_restart(base)
    assert attrs["it"] == 2 and attrs["q"] == list(prm["q"]) and attrs["delt"] == prm["delt"]
    for r in range(2):
        assert np.array_equal(up2[r].view(np.int64), ups[r].view(np.int64))
        assert np.array_equal(np22[r], np2s[r]) and np.array_equal(uf2[r], ufs[r])
    w.close()


@pytest.mark.gpu
def test_restart_into_fresh_context_matches_oracle(tmp_path):
    """Save a device state, load it into a fresh Context (rows re-bucketed on upload, CG warm start zero, as in the
    reference's restart path proj/weibel/app.f90:349-353) and step: same as an oracle world restarted from the file."""
    import wumingpic2d_b200 as wm
    from wumingpic2d_b200 import snapshot as S
    prm, w0 = make_world(40, 24, 8, steps=3)
    s = oracle_state(w0)
    a = wm.Context.from_params(prm)
    a.upload_particles_sorted(s["up"], s["np2"], s["cumcnt"])
    a.upload_field(s["uf"])
    a.step(2)
    base = str(tmp_path / "snap")
    S.save_context(base, a, 5, _cfg(prm))
    a.close()
    attrs, up, np2, uf = S.read_restart(base)
    # oracle restarted from the file: up -> gp, sort__bucket, df = 0
    w = O.World(prm)
    w.array(0, O.GP)[...] = up[0]
    w.array(0, O.NP2)[...] = np2[0]
    w.array(0, O.UF)[...] = uf[0]
    w.sort_bucket()
    b = wm.Context.from_params(prm)
    S.load_into_context(base, b)
    for it in range(2):
        w.step(1); b.step(1)
        assert b.cg_iters() == w.cg_iters()
        upb, np2b, cum = b.download_particles()
        assert np.array_equal(cum, w.array(0, O.CUMCNT))
        x, y = flatten_by_id(upb, np2b), flatten_by_id(w.array(0, O.UP), w.array(0, O.NP2))
        assert np.array_equal(x[0], y[0])
        ex, eu = particle_err(x[2], y[2], prm["nx"], prm["vte"])
        tol = 1e-12 if it == 0 else 1e-10
        assert ex <= tol and eu <= tol
        assert rel_to_max(b.download_field(), w.array(0, O.UF)).max() <= tol
    b.close(); w.close(); w0.close()


def test_moment_file_matches_paraio_arithmetic(tmp_path):
    """io__mom (common/paraio.f90:555-713): density, first and second moments divided by the density, cell-centred
    fields with the reference's own averages (Bz: uf(3,i+1,j) twice, :684), checked against plain loops."""
    from wumingpic2d_b200 import snapshot as S
    prm, w = make_world(12, 8, 4, steps=2)
    w.mom_accl(); w.mom_nvt(); w.bc_mom()
    mom, uf = w.array(0, O.MOM).copy(), w.array(0, O.UF).copy()
    base = str(tmp_path / "0000002_mom")
    root = S.write_mom(base, 2, dict(_cfg(prm), np=prm["np"]), mom, uf)
    assert list(root["dataset"]) == ["den", "vel", "temp", "uf"]
    assert root["dataset"]["vel"]["shape"] == [3, 12, 8, 2] and root["dataset"]["uf"]["shape"] == [6, 12, 8]
    d = S.read_datasets(base)
    assert d["attribute"]["it"] == 2 and d["attribute"]["nsp"] == 2
    den, vel, temp, cc = (d["dataset"][k] for k in ("den", "vel", "temp", "uf"))
    nx, ny = 12, 8
    for isp in range(2):
        for j in range(ny):
            for i in range(nx):
                m = mom[isp, j + 1, i + 1]
                assert den[isp, j, i] == m[0]
                assert np.array_equal(vel[isp, j, i], m[1:4] / m[0]) and np.array_equal(temp[isp, j, i], m[4:7] / m[0])
    for j in range(ny):
        for i in range(nx):
            u = lambda c, di, dj: uf[j + 2 + dj, i + 2 + di, c]
            ref = [(u(0, 0, 0) + u(0, 0, 1)) / 2, (u(1, 0, 0) + u(1, 1, 0)) / 2,
                   (u(2, 0, 0) + u(2, 1, 0) + u(2, 1, 0) + u(2, 1, 1)) / 4,
                   (u(3, 0, 0) + u(3, 1, 0)) / 2, (u(4, 0, 0) + u(4, 0, 1)) / 2, u(5, 0, 0)]
            assert np.array_equal(cc[j, i], ref)
    w.close()


def test_endian_flag_of_a_reference_written_file(tmp_path):
    """A snapshot as the Fortran code writes it on x86 -- meta.endian = 1 (utils/iocore/mpiio.f90:97-109), little-endian
    raw data -- must read back unswapped, and one labelled 16777216 with big-endian raw data must be swapped
    (python/json2hdf5.py:38-41 maps 1 -> '<', 16777216 -> '>')."""
    from wumingpic2d_b200 import snapshot as S
    vals = np.arange(1, 7, dtype=np.float64) * 0.5
    ints = np.array([3, 1, 4], dtype=np.int32)
    for flag, order in ((1, "<"), (16777216, ">")):
        base = str(tmp_path / ("e%d" % flag))
        with open(base + ".raw", "wb") as f:
            f.write(ints.astype(order + "i4").tobytes())
            f.write(vals.astype(order + "f8").tobytes())
        root = {"meta": {"endian": flag, "rawfile": os.path.basename(base) + ".raw"},
                "attribute": {"n": {"datatype": "i4", "offset": 0, "size": 12, "ndim": 1, "shape": [3], "description": "", "data": [3, 1, 4]}},
                "dataset": {"a": {"datatype": "f8", "offset": 12, "size": 48, "ndim": 2, "shape": [3, 2], "description": "a"}}}
        json.dump(root, open(base + ".json", "w"))
        out = S.read_datasets(base)
        assert out["attribute"]["n"] == [3, 1, 4]
        assert np.array_equal(out["dataset"]["a"], vals.reshape(2, 3))
    assert S.endian_flag() == int(np.frombuffer(np.array([1, 0, 0, 0], np.uint8).tobytes(), np.int32)[0])


def test_metadata_writer_reproduces_the_reference_fixture(tmp_path):
    """The one golden file the reference ships for this format: utils/iocore/jsonio_expected.json, which its jsonio_test
    (utils/iocore/jsonio_test.f90: nx=16, ny=32, ns=2, mass, charge, emf (6,ny,nx), mom (10,ny,nx), meta info /
    large_int) must reproduce (utils/iocore/unittest.py:63-76).  The same calls through this package's writer give
    the same JSON key by key -- names, order, datatypes, offsets, sizes, shapes, descriptions, data
    (tests/golden/jsonio_expected.json = that fixture parsed and re-serialised by tests/golden/make_jsonio_fixture.py)."""
    from wumingpic2d_b200 import snapshot as S
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "jsonio_expected.json")))
    nx, ny = 16, 32
    base = str(tmp_path / "test")
    attrs = [("nx", "i4", nx, "# grid in x"), ("ny", "i4", ny, "# grid in y"), ("ns", "i4", 2, "# species"),
             ("mass", "f8", [1.0, 100.0], "mass"), ("charge", "f8", [-1.0, 1.0], "charge")]
    ds = [("emf", "f8", np.zeros((nx, ny, 6)), [6, ny, nx], "electromagnetic fields"),
          ("mom", "f8", np.zeros((nx, ny, 10)), [10, ny, nx], "moments")]
    S._write_file(base, attrs, ds, meta={"info": "some information", "large_int": 2 ** 32})
    mine = json.load(open(base + ".json"))
    if sys.byteorder != "little":
        gold["meta"]["endian"] = 16777216
    assert list(mine) == list(gold)
    for sec in gold:
        assert list(mine[sec]) == list(gold[sec]), sec
        for name, rec in gold[sec].items():
            if isinstance(rec, dict):
                assert list(mine[sec][name]) == list(rec), (sec, name)
            assert mine[sec][name] == rec, (sec, name, mine[sec][name], rec)
    assert os.path.getsize(base + ".raw") == gold["dataset"]["mom"]["offset"] + gold["dataset"]["mom"]["size"]


@pytest.mark.gpu
def test_compare_with_reference_run_script(tmp_path):
    """scripts/compare_with_reference_run.py -- the script a maintainer with gfortran + MPI runs on two restart snapshots of
    the real reference binary -- end to end, with the oracle standing in for the Fortran code: a 2-rank run writes a snapshot
    at step 3, is restarted from it (df = 0, rows re-bucketed, proj/weibel/app.f90:349-353), runs one step and writes the
    second snapshot; the script must load the first, advance the device and find parity with the second."""
    import subprocess
    import sys
    from wumingpic2d_b200 import snapshot as S
    prm, w0 = make_world(24, 16, 6, nranks=2, steps=3)
    cfg = dict(_cfg(prm), np=prm["np"])
    b1 = str(tmp_path / "0000003_restart")
    S.write_restart(b1, 3, prm["nxgs"], prm["nxgs"] + prm["nx"] - 1, cfg, [w0.array(r, O.UP).copy() for r in range(2)],
                    [w0.array(r, O.NP2).copy() for r in range(2)], [w0.array(r, O.UF).copy() for r in range(2)])
    w = O.World(prm)                       # the restarted run
    for r in range(2):
        w.array(r, O.GP)[...] = w0.array(r, O.UP)
        w.array(r, O.NP2)[...] = w0.array(r, O.NP2)
        w.array(r, O.UF)[...] = w0.array(r, O.UF)
    w.sort_bucket()
    w.step(1)
    b2 = str(tmp_path / "0000004_restart")
    S.write_restart(b2, 4, prm["nxgs"], prm["nxgs"] + prm["nx"] - 1, cfg, [w.array(r, O.UP).copy() for r in range(2)],
                    [w.array(r, O.NP2).copy() for r in range(2)], [w.array(r, O.UF).copy() for r in range(2)])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "compare_with_reference_run.py"), b1, b2],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "PARITY WITH THE REFERENCE BINARY: ok" in r.stdout and "bit-exact: True" in r.stdout
    w.close(); w0.close()
