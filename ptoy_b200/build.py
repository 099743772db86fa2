"""In-tree build of the CUDA engine: nvcc -> ptoy_b200/libptoy_b200.so (sm_100a only).

Every translation unit is compiled to an object file in ptoy_b200/build/ (in parallel: the
stage kernels are templates with many instantiations) and linked into one shared library.
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.environ.get("PTOY_B200_LIB", os.path.join(HERE, "libptoy_b200.so"))
NVCC = os.environ.get("PTOY_NVCC", "/usr/local/cuda/bin/nvcc")
# (source, extra defines, object name)
UNITS = [("kernels.cu", [], "kernels"), ("engine.cu", [], "engine"), ("dist.cu", [], "dist"),
         ("c_api.cu", [], "c_api")] + \
        [("stage_tile.cu", [f"-DPTOY_GROUP={g}"], f"stage_tile_g{g}") for g in range(4)]
CFLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function",
    "-ccbin", "/usr/bin/g++",
]
LDFLAGS = ["-shared", "-cudart", "static", "-ccbin", "/usr/bin/g++"]


def _deps():
    d = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    d.append(os.path.join(HERE, "..", "include", "ptoy_b200.h"))
    return d


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def _compile(unit, extra, verbose, tag):
    src, defs, name = unit
    obj = os.path.join(OBJ, f"{name}{tag}.o")
    cmd = [NVCC] + CFLAGS + extra + defs + (["-Xptxas", "-v"] if verbose else []) + \
        ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return obj, r


def check_no_packed_fma(objects):
    """The stage kernels use the packed fp32 forms (sub.rn.f32x2 / mul.rn.f32x2) for differences and
    products only.  ptxas contracts a packed multiply followed by a packed add into FFMA2 - which
    would break the one-rounding-per-reference-operation contract - so no FFMA2
    may appear in the stage kernels' SASS."""
    cuobjdump = os.path.join(os.path.dirname(NVCC), "cuobjdump")
    for o in objects:
        r = subprocess.run([cuobjdump, "-sass", o], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("cuobjdump failed on " + o + ": " + r.stderr[-300:])
        if "FFMA2" in r.stdout:
            raise RuntimeError(o + ": FFMA2 in the stage kernels - a packed multiply-add was contracted")


def build(force=False, verbose=False, lib=None, extra=None):
    """Build `lib` (default: the in-tree library).  `extra`: additional nvcc flags (e.g. -D
    switches of a tuning variant; then pass a different `lib`)."""
    lib = lib or LIB
    if not force and lib == LIB and not stale():
        return lib
    extra = list(extra or []) + os.environ.get("PTOY_NVCC_EXTRA", "").split()
    tag = "" if not extra else "_" + hashlib.md5(" ".join(extra).encode()).hexdigest()[:8]
    os.makedirs(OBJ, exist_ok=True)
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda u: _compile(u, extra, verbose, tag), UNITS))
    for obj, r in results:
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("nvcc failed building " + obj)
        if verbose:
            print(r.stdout + r.stderr)
    check_no_packed_fma([o for o, _ in results if "stage_tile" in os.path.basename(o)])
    r = subprocess.run([NVCC] + LDFLAGS + ["-o", lib] + [o for o, _ in results] + ["-ldl"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("linking libptoy_b200.so failed")
    return lib


if __name__ == "__main__":
    build(force=True, verbose="-v" in sys.argv)
    print(LIB)
