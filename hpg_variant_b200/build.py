"""Build the CUDA library in-tree (hpg_variant_b200/libhpgv_epi.so) for sm_100a.

nvcc cross-compiles without a GPU.  The built .so is git-ignored but travels to
the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libhpgv_epi.so")
HOSTLIB = os.path.join(HERE, "libhpgv_epi_host.so")
CLI = os.path.join(HERE, "hpg-var-gwas-b200")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-ccbin", "g++",
]
# one translation unit per kernel family (compiled in parallel: the search kernels dominate the build) + the C-ABI
CUDA_UNITS = ["epi_capi.cu", "epi_k_search2.cu", "epi_k_search2_tri.cu", "epi_k_search3.cu", "epi_k_search3v2.cu", "epi_k_search3v3.cu"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _env():
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    return env


def build(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h"))]
    inc = [os.path.join(os.path.dirname(HERE), "include", "hpgv_epi.h")]
    if force or _newer(LIB, srcs + inc):
        from concurrent.futures import ThreadPoolExecutor
        objdir = os.path.join(HERE, "build")
        os.makedirs(objdir, exist_ok=True)

        def compile_unit(name):
            obj = os.path.join(objdir, name[:-3] + ".o")
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, name)]
            if verbose:
                print(" ".join(cmd), file=sys.stderr)
            r = subprocess.run(cmd, env=_env(), capture_output=True, text=True)
            return name, obj, r

        with ThreadPoolExecutor(max_workers=len(CUDA_UNITS)) as pool:
            results = list(pool.map(compile_unit, CUDA_UNITS))
        for name, obj, r in results:
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise subprocess.CalledProcessError(r.returncode, "nvcc " + name)
        subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "g++", "-o", LIB] + [obj for _, obj, _ in results],
                       check=True, env=_env())
    host_src = os.path.join(CSRC, "epi_host.cpp")
    if os.path.exists(host_src):
        compat = os.path.join(os.path.dirname(HERE), "include", "hpgv_epi_compat.h")
        if force or _newer(HOSTLIB, [host_src, compat] + inc + [LIB]):
            # no CUDA runtime here: everything that touches the GPU goes through the C-ABI of libhpgv_epi.so
            subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-o", HOSTLIB, host_src,
                            "-L" + HERE, "-lhpgv_epi", "-Wl,-rpath,$ORIGIN"], check=True, env=_env())
        cli_src = os.path.join(CSRC, "epi_cli.cpp")
        if os.path.exists(cli_src) and (force or _newer(CLI, [cli_src, HOSTLIB])):
            subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", CLI, cli_src, "-L" + HERE, "-lhpgv_epi_host", "-lhpgv_epi",
                            "-Wl,-rpath,$ORIGIN"], check=True, env=_env())
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
