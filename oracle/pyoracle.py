"""ctypes binding of oracle/_build/libxsmm_oracle.so (TEST INFRASTRUCTURE).

Argument names and order mirror the C-ABI of the product (include/tpp_xsmm_abi.h)
so the parity tests read ``oracle.brgemm(...)`` next to ``xsmm.brgemm(...)``.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import c_int, c_int64, c_void_p

import numpy as np

F32 = 1
BF16 = 2

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_lib_dir = None
_native = False


def use_native(flag: bool = True) -> None:
    """Select the -march=native build (oracle/_build_native) for every later call. bench.py's CPU arm uses it on
    the box it is timed on; tests use the portable x86-64-v3 build that travels with the repo."""
    global _native
    _native = bool(flag)


def build(native: bool = False, force: bool = False) -> str:
    """Compile the oracle with the Makefile next to this file; returns the .so path.

    native=True builds with -march=native into oracle/_build_native (used by
    bench.py for the CPU arm on the box it is timed on); the default x86-64-v3
    build is the one that travels.
    """
    out = "_build_native" if native else "_build"
    so = os.path.join(_HERE, out, "libxsmm_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("xsmm_oracle.c", "xsmm_oracle_fast.c", "xsmm_oracle.h", "tensor_init.cpp",
                                             "Makefile")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        cmd = ["make", "-C", _HERE, f"OUT={out}"]
        if native:
            cmd.append("ARCH=native")
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    return so


def lib(native: bool | None = None):
    global _lib, _lib_dir
    native = _native if native is None else native
    want = "_build_native" if native else "_build"
    if _lib is not None and _lib_dir == want:
        return _lib
    L = ctypes.CDLL(build(native=native))
    i64 = c_int64
    L.xo_brgemm.argtypes = [i64] * 10 + [c_void_p, c_void_p, c_void_p, i64]
    L.xo_brgemm.restype = None
    L.xo_gemm.argtypes = [i64] * 8 + [c_void_p, c_void_p, c_void_p]
    L.xo_gemm.restype = None
    L.xo_fused_brgemm.argtypes = [i64] * 14 + [c_void_p] * 4 + [i64]
    L.xo_fused_brgemm.restype = None
    L.xo_unary.argtypes = [i64] * 7 + [c_void_p, c_void_p]
    L.xo_unary.restype = c_int
    L.xo_binary.argtypes = [i64] * 8 + [c_void_p, c_void_p, c_void_p]
    L.xo_binary.restype = c_int
    L.xo_f32_to_bf16_array.argtypes = [c_void_p, c_void_p, i64]
    L.xo_f32_to_bf16_array.restype = None
    L.xo_bf16_to_f32_array.argtypes = [c_void_p, c_void_p, i64]
    L.xo_bf16_to_f32_array.restype = None
    L.xo_set_acc_mode.argtypes = [c_int]
    L.xo_set_num_threads.argtypes = [c_int]
    L.xo_num_threads.restype = c_int
    L.xo_set_vnni_factor.argtypes = [c_int]
    L.xo_fused_brgemm_fast.argtypes = [i64] * 13 + [c_void_p] * 4 + [i64]
    L.xo_fused_brgemm_fast.restype = c_int
    L.xo_fast_isa.restype = c_int
    L.xo_fused_brgemm_amx.argtypes = [i64] * 13 + [c_void_p] * 4 + [i64]
    L.xo_fused_brgemm_amx.restype = c_int
    L.xo_fused_brgemm_amx_grid.argtypes = [i64] * 13 + [c_void_p] * 4 + [i64] * 8
    L.xo_fused_brgemm_amx_grid.restype = c_int
    L.ti_create.argtypes = [c_int, c_int, c_int]
    L.ti_create.restype = c_void_p
    L.ti_destroy.argtypes = [c_void_p]
    L.ti_fill.argtypes = [c_void_p, i64, c_void_p]
    L.ti_fill.restype = None
    _lib, _lib_dir = L, want
    return L


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def np_dtype(dtype: int):
    return np.float32 if dtype == F32 else np.uint16


def f32_to_bf16(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    out = np.empty(a.shape, dtype=np.uint16)
    lib().xo_f32_to_bf16_array(_p(a), _p(out), a.size)
    return out


def bf16_to_f32(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint16)
    out = np.empty(a.shape, dtype=np.float32)
    lib().xo_bf16_to_f32_array(_p(a), _p(out), a.size)
    return out


def set_acc_mode(mode: int) -> None:
    lib().xo_set_acc_mode(mode)


def set_num_threads(n: int) -> None:
    lib().xo_set_num_threads(n)


def set_vnni_factor(v: int) -> None:
    """VNNI blocking factor of B operands carrying gemm flag 2048: 2 (default) or 4 (mlir-gen --vnni=4 layouts)."""
    lib().xo_set_vnni_factor(v)


def num_threads() -> int:
    return lib().xo_num_threads()


def brgemm(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, flags, A, B, C, batch):
    lib().xo_brgemm(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, flags, _p(A), _p(B), _p(C), batch)


def gemm(dtype, m, n, k, lda, ldb, ldc, flags, A, B, C):
    lib().xo_gemm(dtype, m, n, k, lda, ldb, ldc, flags, _p(A), _p(B), _p(C))


def fused_brgemm(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, gemm_flags, unary_flags, unary_kind,
                 binary_flags, binary_kind, A, B, C, D, batch):
    lib().xo_fused_brgemm(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, gemm_flags, unary_flags, unary_kind,
                          binary_flags, binary_kind, _p(A), _p(B), _p(C), _p(D), batch)


def fused_brgemm_fast(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, gemm_flags, unary_kind, binary_flags,
                      binary_kind, A, B, C, D, batch) -> bool:
    """Vectorised CPU kernel for the timed CPU arm (xsmm_oracle_fast.c). Returns False if the shape is not
    supported (caller falls back to the plain oracle)."""
    return lib().xo_fused_brgemm_fast(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, gemm_flags, unary_kind,
                                      binary_flags, binary_kind, _p(A), _p(B), _p(C), _p(D), batch) == 0


def fused_brgemm_amx(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, gemm_flags, unary_kind, binary_flags,
                     binary_kind, A, B_vnni2, C, D, batch) -> bool:
    """AMX-BF16 tile kernel (tdpbf16ps) for a VNNI-2 packed B (gemm flag 2048) - the instruction and the weight layout
    libxsmm's JIT uses on Sapphire Rapids and later. False when the build / host / kernel has no AMX or the shape is not
    a multiple of 32 (caller falls back)."""
    return lib().xo_fused_brgemm_amx(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, gemm_flags, unary_kind,
                                     binary_flags, binary_kind, _p(A), _p(B_vnni2), _p(C), _p(D), batch) == 0


def fused_brgemm_amx_grid(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, gemm_flags, unary_kind, binary_flags,
                          binary_kind, A, B_vnni2, C, D, batch, grid_n, grid_k, a_step, b_step, c_step_n, c_step_k,
                          d_step) -> bool:
    """A whole layer as the reference runs it: grid_n x grid_k tile BRGEMMs (AMX-BF16) on block-packed operands, the
    (iN, iK) loop nest shared by the OpenMP threads (scf.parallel in the reference)."""
    return lib().xo_fused_brgemm_amx_grid(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, gemm_flags, unary_kind,
                                          binary_flags, binary_kind, _p(A), _p(B_vnni2), _p(C), _p(D), batch, grid_n,
                                          grid_k, a_step, b_step, c_step_n, c_step_k, d_step) == 0


def fast_isa() -> str:
    return {3: "AMX-BF16 tdpbf16ps 32x32 tile block", 2: "AVX512-BF16 vdpbf16ps 8x32 microkernel",
            1: "compiler-vectorised f32 FMA"}[lib().xo_fast_isa()]


def has_amx() -> bool:
    return lib().xo_fast_isa() == 3


def unary(kind, dtype, m, n, ldi, ldo, flags, inp, out):
    rc = lib().xo_unary(kind, dtype, m, n, ldi, ldo, flags, _p(inp), _p(out))
    if rc != 0:
        raise ValueError(f"oracle: unsupported unary kind={kind} dtype={dtype} m={m}")


def binary(kind, dtype, m, n, ldl, ldr, ldo, flags, lhs, rhs, out):
    rc = lib().xo_binary(kind, dtype, m, n, ldl, ldr, ldo, flags, _p(lhs), _p(rhs), _p(out))
    if rc != 0:
        raise ValueError(f"oracle: unsupported binary kind={kind}")


class TensorInit:
    """One generator per (type, dtype, seed), filled sequentially, like
    getTensorInit() (lib/TPP/Transforms/Utils/TensorInit.cpp:75-144)."""

    TYPES = {"const": 0, "simple": 1, "cont": 2, "random": 3, "normal": 4}

    def __init__(self, kind: str, dtype: int, seed: int = 0):
        self.dtype = dtype
        self._h = lib().ti_create(self.TYPES[kind], dtype, seed)

    def fill(self, *shape) -> np.ndarray:
        out = np.empty(shape, dtype=np_dtype(self.dtype))
        lib().ti_fill(self._h, out.size, _p(out))
        return out

    def __del__(self):
        try:
            lib().ti_destroy(self._h)
        except Exception:
            pass


# ---- tensor.pack / tensor.unpack (block layouts) ------------------------------------------------------
# Restates the semantics of the ops the reference tiles and lowers to per-tile unary TPPs
# (lib/TPP/Transforms/LowerPacksAndUnpacks.cpp:143-250; benchmarks/mlir/fp32-pack-gemm-operand-{a,b}-512x1024.mlir,
# fp32-unpack-gemm-operand-a-512x512.mlir): inner_dims_pos = [0, 1], inner_tiles = [bm, bn], optional
# outer_dims_perm = [1, 0]. Pure data movement: bit-exact for every dtype.
def tensor_pack(x, bm, bn, outer_perm=(0, 1)):
    """[M][N] -> [M/bm][N/bn][bm][bn] (outer dims permuted by outer_perm)."""
    import numpy as np

    m, n = x.shape
    assert m % bm == 0 and n % bn == 0
    t = x.reshape(m // bm, bm, n // bn, bn).transpose(0, 2, 1, 3)
    if tuple(outer_perm) == (1, 0):
        t = t.transpose(1, 0, 2, 3)
    return np.ascontiguousarray(t)


def tensor_unpack(xp, outer_perm=(0, 1)):
    """[A][B][bm][bn] -> [M][N], the inverse of tensor_pack."""
    import numpy as np

    t = xp
    if tuple(outer_perm) == (1, 0):
        t = t.transpose(1, 0, 2, 3)
    mb, nb, bm, bn = t.shape
    return np.ascontiguousarray(t.transpose(0, 2, 1, 3).reshape(mb * bm, nb * bn))
