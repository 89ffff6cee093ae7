"""Build libemplanner.so in-tree with nvcc for sm_100a (no JIT, no torch extension machinery).

sl_dp.cu is compiled with -fmad=false: its arithmetic must match the float64
oracle bit for bit (numpy never contracts a*b+c).  The other translation units
keep FMA contraction on.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libemplanner.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
          "--expt-relaxed-constexpr"]
UNITS = [("sl_dp.cu", ["-fmad=false"]), ("sl_qp.cu", []), ("st_dp.cu", ["-fmad=false"]), ("frenet.cu", ["-fmad=false"]), ("smooth_qp.cu", []), ("misc.cu", ["-fmad=false"]), ("st_misc.cu", ["-fmad=false"]),
         ("st_back.cu", ["-fmad=false"]), ("st_qp.cu", []), ("cycle.cu", ["-fmad=false"]),
         ("fp64_peak.cu", []), ("capi.cu", [])]


def _nvcc():
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libemplanner.so cannot be built")
    return cand


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    nvcc = _nvcc()
    defines = os.environ.get("EMP_DEFINES", "").split()      # tuning experiments, e.g. EMP_DEFINES="-DEMP_QP_MINBLOCKS=3"
    if defines:
        force = True
    hdrs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(os.path.dirname(PKG), "include", "emplanner.h"))
    hdrs.append(os.path.abspath(__file__))
    objs = []
    for src, extra in UNITS:
        path = os.path.join(HERE, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(HERE, src.replace(".cu", ".o"))
        if force or _stale(obj, [path] + hdrs):
            cmd = [nvcc] + ARCH + COMMON + extra + defines + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.run(cmd, check=True)
        objs.append(obj)
    if force or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
