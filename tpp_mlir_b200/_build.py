"""In-tree nvcc build of the drop-in library ``libtpp_xsmm_runner_utils.so``.

The reference builds a library of the same name from runtime/Xsmm + runtime/
PerfRunnerUtils.cpp against libxsmm (runtime/Xsmm/CMakeLists.txt:1-11); this one
is built from tpp_mlir_b200/csrc for sm_100a only. nvcc cross-compiles without a
GPU, so this runs on the CPU-only build box; the resulting .so travels with the
source tree to the GPU box.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIBNAME = "libtpp_xsmm_runner_utils.so"

SOURCES = ["runtime.cu", "eltwise.cu", "brgemm_simt.cu", "tc_host.cu", "brgemm_tc.cu", "mlp_chain.cu", "mlp_chain_ft.cu",
           "mlp_chain_pair.cu", "tile_grid.cu", "vnni_flat.cu"]
HEADERS = ["common.cuh", "kernels.h", "kernel_desc.h", "ptx.cuh", "tc_common.cuh", "tc_splitk.cuh", os.path.join(ROOT, "include", "tpp_xsmm_abi.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; the CUDA extension cannot be built")


def lib_path() -> str:
    return os.path.join(LIBDIR, LIBNAME)


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link the shared library. Returns its path."""
    os.makedirs(OBJDIR, exist_ok=True)
    hdrs = [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    procs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + hdrs + [os.path.abspath(__file__)]):
            cmd = [nvcc()] + NVCC_FLAGS + ["-c", s, "-o", o]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        with open(os.path.join(OBJDIR, src + ".ptxas.log"), "w") as f:
            f.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    so = lib_path()
    if force or procs or _stale(so, objs):
        cmd = [nvcc(), "-shared", "-o", so] + objs + ["-cudart", "static", "-Xcompiler", "-fPIC"]
        subprocess.run(cmd, check=True)
    return so


REPLAY_NAME = "libtpp_replay.so"


def replay_path() -> str:
    return os.path.join(LIBDIR, REPLAY_NAME)


def build_replay(force: bool = False) -> str:
    """The native tpp-run stand-in loop (csrc/harness/replay.cpp): plain C++, calls only the C-ABI."""
    src = os.path.join(CSRC, "harness", "replay.cpp")
    so = replay_path()
    lib = build()
    if force or _stale(so, [src, lib, os.path.join(ROOT, "include", "tpp_xsmm_abi.h")]):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-I", os.path.join(ROOT, "include"),
               src, "-o", so, "-L", LIBDIR, "-ltpp_xsmm_runner_utils", "-Wl,-rpath,$ORIGIN"]
        subprocess.run(cmd, check=True)
    return so


STANDIN_NAME = "tpp_run_standin"


def standin_path() -> str:
    return os.path.join(LIBDIR, STANDIN_NAME)


def build_standin(force: bool = False) -> str:
    """tpp_run_standin (csrc/harness/tpp_run_standin.cpp): what `mlir-gen ... | tpp-run -n N` executes, as a native
    program that links only the C-ABI library."""
    src = os.path.join(CSRC, "harness", "tpp_run_standin.cpp")
    exe = standin_path()
    lib = build()
    if force or _stale(exe, [src, lib, os.path.join(ROOT, "include", "tpp_xsmm_abi.h")]):
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        cmd = [cxx, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", LIBDIR,
               "-ltpp_xsmm_runner_utils", "-Wl,-rpath,$ORIGIN"]
        subprocess.run(cmd, check=True)
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_replay(force="--force" in sys.argv))
    print(build_standin(force="--force" in sys.argv))
