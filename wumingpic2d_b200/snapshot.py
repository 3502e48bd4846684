"""Restart snapshots in the reference's format: the pair NAME.json + NAME.raw written by
`paraio__output` (common/paraio.f90:93-228) and read back by `paraio__input` (:233-426).

The raw file is a flat concatenation in native byte order; the JSON gives names, types, byte offsets and column-major
shapes (utils/iocore/jsonio.f90:128-192).  Attributes, in this order (paraio.f90:143-155, 963-1000):
    dummy_attribute (i4, = MPI_OFFSET_KIND), it, nxs, nxe, ndim, np, nxgs, nxge, nygs, nyge, nsp, nproc (i4),
    delx, delt, c (f8), r(nsp), q(nsp) (f8)
Datasets (paraio.f90:168-213), every rank's whole local array, rank-major:
    up  f8 [ndim, np, nyl, nsp, nproc]      np2 i4 [nyl, nsp, nproc]      uf  f8 [6, nx+4, nyl+4, nproc]  (with ghosts)

Arrays on the Python side are in C order = reversed Fortran order, like everywhere in this package
(`up[rank].shape == (nsp, nyl, np, 6)`, `np2[rank].shape == (nsp, nyl)`, `uf[rank].shape == (nyl+4, nx+4, 6)`), so a
dataset is simply the concatenation of the ranks' buffers.

A snapshot written by a real reference run (built elsewhere) can be loaded as the initial condition of a Context and
its next snapshot compared: the only route to parity against the actual Fortran binary.
"""
import json
import os
import sys

import numpy as np

_DT = {"i4": np.int32, "i8": np.int64, "f4": np.float32, "f8": np.float64}
_I4_ATTRS = ("dummy_attribute", "it", "nxs", "nxe", "ndim", "np", "nxgs", "nxge", "nygs", "nyge", "nsp", "nproc")
_F8_ATTRS = ("delx", "delt", "c", "r", "q")


def endian_flag():
    """mpiio_get_endian_flag (utils/iocore/mpiio.f90:97-109): the bytes (1, 0, 0, 0) transferred into an int32, i.e. 1
    on a little-endian writer and 16777216 on a big-endian one; python/json2hdf5.py:38-41 reads them as '<' and '>'."""
    return 1 if sys.byteorder == "little" else 16777216


def write_restart(base, it, nxs, nxe, cfg, up, np2, uf):
    """Write `base`.json + `base`.raw.  cfg: dict with ndim, np, nxgs, nxge, nygs, nyge, nsp, delx, delt, c, r, q.
    up, np2, uf: lists with one array per rank (equal nyl on every rank, as the format requires)."""
    nproc = len(up)
    nsp, nyl, npcap, ndim = up[0].shape
    if any(a.shape != up[0].shape for a in up) or any(a.shape != uf[0].shape for a in uf):
        raise ValueError("all ranks must hold the same number of rows (common/paraio.f90:168-213)")
    vals = dict(dummy_attribute=8, it=it, nxs=nxs, nxe=nxe, ndim=ndim, np=npcap, nxgs=cfg["nxgs"], nxge=cfg["nxge"],
                nygs=cfg["nygs"], nyge=cfg["nyge"], nsp=nsp, nproc=nproc, delx=cfg["delx"], delt=cfg["delt"], c=cfg["c"],
                r=list(cfg["r"])[:nsp], q=list(cfg["q"])[:nsp])
    root = {"meta": {"endian": endian_flag(), "rawfile": os.path.basename(base) + ".raw"}, "attribute": {}, "dataset": {}}
    disp = 0
    with open(base + ".raw", "wb") as f:
        for name in _I4_ATTRS + _F8_ATTRS:
            dt = "i4" if name in _I4_ATTRS else "f8"
            a = np.atleast_1d(np.asarray(vals[name], dtype=_DT[dt]))
            root["attribute"][name] = {"datatype": dt, "offset": disp, "size": int(a.nbytes), "ndim": 1, "shape": [int(a.size)],
                                       "description": "", "data": (a.tolist() if a.size > 1 else a.tolist()[0])}
            f.write(a.tobytes())
            disp += a.nbytes
        nxg = uf[0].shape[1]
        for name, dt, arrs, shape, desc in (
                ("up", "f8", up, [ndim, npcap, nyl, nsp, nproc], "particles"),
                ("np2", "i4", np2, [nyl, nsp, nproc], "number of active particles"),
                ("uf", "f8", uf, [6, nxg, uf[0].shape[0], nproc], "electromagnetic fields including ghost cells")):
            size = 0
            for a in arrs:
                b = np.ascontiguousarray(a, dtype=_DT[dt])
                f.write(b.tobytes())
                size += b.nbytes
            root["dataset"][name] = {"datatype": dt, "offset": disp, "size": int(size), "ndim": len(shape), "shape": shape,
                                     "description": desc}
            disp += size
    with open(base + ".json", "w") as f:
        json.dump(root, f, indent=2)
    return root


def read_restart(base):
    """Read `base`.json + its raw file.  Returns (attrs: dict, up, np2, uf: lists with one array per rank)."""
    with open(base + ".json") as f:
        root = json.load(f)
    swap = root["meta"]["endian"] != endian_flag()
    raw = os.path.join(os.path.dirname(base), root["meta"]["rawfile"])

    def fetch(rec):
        dt = np.dtype(_DT[rec["datatype"]])
        a = np.fromfile(raw, dtype=dt, count=rec["size"] // dt.itemsize, offset=rec["offset"])
        return a.byteswap() if swap else a

    attrs = {}
    for name, rec in root["attribute"].items():
        a = fetch(rec)
        attrs[name] = a.tolist() if a.size > 1 else a.tolist()[0]
    ds = root["dataset"]
    ndim, npcap, nyl, nsp, nproc = ds["up"]["shape"]
    up = list(fetch(ds["up"]).reshape(nproc, nsp, nyl, npcap, ndim))
    np2 = list(fetch(ds["np2"]).reshape(nproc, nsp, nyl))
    six, nxg, nylg, _ = ds["uf"]["shape"]
    uf = list(fetch(ds["uf"]).reshape(nproc, nylg, nxg, six))
    return attrs, up, np2, uf


def save_context(base, ctx, it, cfg, comm_gather=None):
    """Snapshot of one Context (one rank; pass lists gathered from all ranks through write_restart for more)."""
    up, np2, _ = ctx.download_particles()
    uf = ctx.download_field()
    return write_restart(base, it, cfg["nxgs"], cfg["nxge"], cfg, [up], [np2], [uf])


def load_into_context(base, ctx, rank=0):
    """paraio__input followed by sort__bucket, as the applications do on restart (proj/weibel/app.f90:349-353): the
    rows of `up` are re-bucketed on upload, the CG warm start starts from zero (the reference's SAVEd df is not part
    of a snapshot either, common/field.f90:98)."""
    attrs, up, np2, uf = read_restart(base)
    ctx.upload_particles(np.ascontiguousarray(up[rank]), np.ascontiguousarray(np2[rank]))
    ctx.upload_field(np.ascontiguousarray(uf[rank]))
    return attrs


def _write_file(base, attrs, datasets, meta=None):
    """Common writer: attrs = [(name, dtype, value[, description])], datasets = [(name, dtype, array(C order), Fortran
    shape, desc)]; meta = extra entries of the "meta" object (jsonio_put_metadata of scalars, utils/iocore/jsonio.f90)."""
    root = {"meta": {"endian": endian_flag(), "rawfile": os.path.basename(base) + ".raw"}, "attribute": {}, "dataset": {}}
    root["meta"].update(meta or {})
    disp = 0
    with open(base + ".raw", "wb") as f:
        for name, dt, val, *desc in attrs:
            a = np.atleast_1d(np.asarray(val, dtype=_DT[dt]))
            root["attribute"][name] = {"datatype": dt, "offset": disp, "size": int(a.nbytes), "ndim": 1, "shape": [int(a.size)],
                                       "description": desc[0] if desc else "", "data": (a.tolist() if a.size > 1 else a.tolist()[0])}
            f.write(a.tobytes())
            disp += a.nbytes
        for name, dt, arr, shape, desc in datasets:
            b = np.ascontiguousarray(arr, dtype=_DT[dt])
            f.write(b.tobytes())
            root["dataset"][name] = {"datatype": dt, "offset": disp, "size": int(b.nbytes), "ndim": len(shape),
                                     "shape": [int(x) for x in shape], "description": desc}
            disp += b.nbytes
    with open(base + ".json", "w") as f:
        json.dump(root, f, indent=2)
    return root


def moment_datasets(mom, uf):
    """The arithmetic of paraio__mom (common/paraio.f90:640-694) on one slab or on the whole domain:
    mom (nsp, nyl+2, nx+2, 7) with one ghost cell per side as mom_calc__nvt + bc__mom leave it, uf (nyl+4, nx+4, 6).
    Returns den (nsp, ny, nx), vel, temp (nsp, ny, nx, 3) -- second moments divided by the density, not centred --
    and the cell-centred fields (ny, nx, 6), literally as the reference averages them (its Bz average takes
    uf(3,i+1,j) twice and never uf(3,i,j+1), paraio.f90:684)."""
    m = np.asarray(mom)[:, 1:-1, 1:-1, :]     # nxgs:nxge, nys:nye
    den = m[..., 0].copy()
    with np.errstate(divide="ignore", invalid="ignore"):
        vel = m[..., 1:4] / m[..., 0:1]
        temp = m[..., 4:7] / m[..., 0:1]
    u = np.asarray(uf)
    c = u[2:-2, 2:-2]
    xp = u[2:-2, 3:-1]      # i+1
    yp = u[3:-1, 2:-2]      # j+1
    xyp = u[3:-1, 3:-1]     # i+1, j+1
    cc = np.empty(c.shape)
    cc[..., 0] = (c[..., 0] + yp[..., 0]) / 2
    cc[..., 1] = (c[..., 1] + xp[..., 1]) / 2
    cc[..., 2] = (c[..., 2] + xp[..., 2] + xp[..., 2] + xyp[..., 2]) / 4
    cc[..., 3] = (c[..., 3] + xp[..., 3]) / 2
    cc[..., 4] = (c[..., 4] + yp[..., 4]) / 2
    cc[..., 5] = c[..., 5]
    return den, vel, temp, cc


def write_mom(base, it, cfg, mom, uf, nproc=1):
    """NNNNNNN_mom.json/.raw as paraio__mom writes it (common/paraio.f90:555-713): attributes dummy_attribute, it and
    the common metadata, datasets den [nx,ny,nsp], vel [3,nx,ny,nsp], temp [3,nx,ny,nsp], uf [6,nx,ny] (global
    arrays; pass the slabs concatenated along y, ghost rows only at the outer ends)."""
    den, vel, temp, cc = moment_datasets(mom, uf)
    nsp, ny, nx = den.shape
    attrs = [("dummy_attribute", "i4", 8), ("it", "i4", it), ("ndim", "i4", 6), ("np", "i4", cfg.get("np", 0)),
             ("nxgs", "i4", cfg["nxgs"]), ("nxge", "i4", cfg["nxge"]), ("nygs", "i4", cfg["nygs"]), ("nyge", "i4", cfg["nyge"]),
             ("nsp", "i4", nsp), ("nproc", "i4", nproc), ("delx", "f8", cfg["delx"]), ("delt", "f8", cfg["delt"]),
             ("c", "f8", cfg["c"]), ("r", "f8", list(cfg["r"])[:nsp]), ("q", "f8", list(cfg["q"])[:nsp])]
    ds = [("den", "f8", den, [nx, ny, nsp], "density"), ("vel", "f8", vel, [3, nx, ny, nsp], "velocity"),
          ("temp", "f8", temp, [3, nx, ny, nsp], "temperature"), ("uf", "f8", cc, [6, nx, ny], "electromagnetic field")]
    return _write_file(base, attrs, ds)


def read_datasets(base):
    """Generic reader: every attribute and dataset of a NAME.json + raw pair; datasets come back in C order
    (reversed Fortran shape)."""
    with open(base + ".json") as f:
        root = json.load(f)
    swap = root["meta"]["endian"] != endian_flag()
    raw = os.path.join(os.path.dirname(base), root["meta"]["rawfile"])
    out = {"attribute": {}, "dataset": {}}
    for kind in ("attribute", "dataset"):
        for name, rec in root[kind].items():
            dt = np.dtype(_DT[rec["datatype"]])
            a = np.fromfile(raw, dtype=dt, count=rec["size"] // dt.itemsize, offset=rec["offset"])
            a = a.byteswap() if swap else a
            if kind == "attribute":
                out[kind][name] = a.tolist() if a.size > 1 else a.tolist()[0]
            else:
                out[kind][name] = a.reshape(rec["shape"][::-1])
    return out
