"""Builds libwumingpic2d.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m wumingpic2d_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels with the tree.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libwumingpic2d.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# field kernels keep the reference's unfused operation order; particle kernels use FMA.  k_fused_sm sits at the 168-register cap,
# where ptxas' register-usage level moves the kernel time (r02at: level 0 +2.0 %, default 5 = reference, 7 / 10 -0.7 %)
UNITS = [("fused_kernel.cu", []), ("fused5_kernel.cu", ["-Xptxas", "--register-usage-level=7"]), ("fused6_kernel.cu", []), ("particle_kernels.cu", []), ("gen_kernels.cu", []), ("hostpipe_kernels.cu", []), ("field_kernels.cu", ["-fmad=false"]), ("cg_persist_kernel.cu", ["-fmad=false"]), ("wm_api.cu", [])]


def source_hash():
    """sha256 over the sources the library is built from (csrc/*.cu, csrc/*.h, include/wumingpic2d.h), first 16 hex digits.
    It is compiled into the library (wm_source_hash) and checked by load_library(): the .so is git-ignored and travels with
    the tree, so a stale binary must fail loudly instead of silently testing yesterday's kernels."""
    import hashlib
    h = hashlib.sha256()
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".h")) and f != "wm_source_hash.h")
    files.append(os.path.join(os.path.dirname(HERE), "include", "wumingpic2d.h"))
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    stamp = os.path.join(CSRC, "wm_source_hash.h")  # generated, git-ignored: only wm_api.cu includes it
    line = '#define WM_SOURCE_HASH "%s"\n' % source_hash()
    if not os.path.exists(stamp) or open(stamp).read() != line:
        with open(stamp, "w") as f:
            f.write(line)
    hdrs = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith(".h") and h != "wm_source_hash.h"]
    hdrs.append(os.path.join(os.path.dirname(HERE), "include", "wumingpic2d.h"))
    objs, rebuilt = [], False
    os.makedirs(os.path.join(HERE, "_obj"), exist_ok=True)
    for src, extra in UNITS:
        s = os.path.join(CSRC, src)
        o = os.path.join(HERE, "_obj", src.replace(".cu", ".o"))
        if force or _newer(s, o) or any(_newer(h, o) for h in hdrs) or (src == "wm_api.cu" and _newer(stamp, o)):
            cmd = [nvcc] + ARCH + COMMON + extra + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            with open(o + ".log", "w") as f:
                f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError("nvcc failed on " + src)
            if verbose:
                sys.stderr.write(r.stderr)
            rebuilt = True
        objs.append(o)
    if rebuilt or not os.path.exists(OUT):
        cmd = [nvcc] + ARCH + ["-shared", "-o", OUT] + objs + ["-lnccl", "-lcudart"]
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
