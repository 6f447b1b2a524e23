"""Python mirror of the ``xsmm`` dialect's dispatch/invoke surface over the C-ABI.

Each function is a 1:1 ctypes call into ``libtpp_xsmm_runner_utils.so`` with the
argument order of the lowered ``func.call`` (ConvertXsmmToFunc.cpp:37-101,
298-352; pinned by test/Conversion/XsmmToFunc/xsmm-to-func.mlir), so parity tests
read like the reference's own IR:

    h = xsmm.brgemm_dispatch(BF16, m, n, k, lda, ldb, ldc, stride_a, stride_b, flags)
    xsmm.brgemm_invoke(BF16, h, A, 0, B, 0, C, 0, batch)

Memref operands are ``(aligned pointer, offset in elements)`` pairs exactly as in
the ABI; a torch tensor (CUDA or CPU), a numpy array or a raw integer address is
accepted for the pointer. There is no fallback: if the CUDA library cannot be
loaded, importing this module raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_void_p

from . import _build

# enum values (include/TPP/Dialect/Xsmm/XsmmEnum.td:13-84)
F32, BF16 = 1, 2
BINARY_NONE, BINARY_ADD, BINARY_MUL, BINARY_SUB, BINARY_DIV = 0, 1, 2, 3, 4
UNARY_NONE, UNARY_IDENTITY, UNARY_ZERO, UNARY_RELU, UNARY_VNNI2, UNARY_TRANSPOSE = 0, 1, 2, 5, 28, 29
UNARY_UNVNNI2_EXT = 1028
UNARY_VNNI4, UNARY_UNVNNI4_EXT = 32, 1032   # [K][N] <-> [K/4][N][4]
UNARY_FLAG_NONE, UNARY_FLAG_BCAST_ROW, UNARY_FLAG_BCAST_COL, UNARY_FLAG_BCAST_SCALAR = 0, 2, 4, 8
BINARY_FLAG_NONE = 0
BINARY_FLAG_BCAST_ROW_IN_0, BINARY_FLAG_BCAST_ROW_IN_1 = 1, 2
BINARY_FLAG_BCAST_COL_IN_0, BINARY_FLAG_BCAST_COL_IN_1 = 4, 8
BINARY_FLAG_BCAST_SCALAR_IN_0, BINARY_FLAG_BCAST_SCALAR_IN_1 = 16, 32
# GEMM flags as the C-ABI receives them (after the lowering's A<->B swap)
GEMM_FLAG_NONE, GEMM_FLAG_BETA_0 = 0, 4
GEMM_FLAG_NO_RESET_TILECONFIG, GEMM_FLAG_NO_SETUP_TILECONFIG = 64, 128
GEMM_FLAG_ROWMAJOR_B_VNNI, GEMM_FLAG_ROWMAJOR_A_VNNI, GEMM_FLAG_VNNI_C = 2048, 4096, 8192

EXPORTS = {
    # name: (restype, argtypes)
    "xsmm_gemm_dispatch": (c_int64, [c_int64] * 8),
    "xsmm_unary_dispatch": (c_int64, [c_int64] * 7),
    "xsmm_binary_dispatch": (c_int64, [c_int64] * 8),
    "xsmm_brgemm_dispatch": (c_int64, [c_int64] * 10),
    "xsmm_fused_brgemm_dispatch": (c_int64, [c_int64] * 14),
    "xsmm_intel_amx_tile_config_dispatch": (c_int64, [c_int64] * 10),
    "xsmm_gemm_invoke": (None, [c_int64, c_int64] + [c_void_p, c_int64] * 3),
    "xsmm_unary_invoke": (None, [c_int64, c_int64] + [c_void_p, c_int64] * 2),
    "xsmm_unary_scalar_invoke": (None, [c_int64, c_int64, c_float, c_void_p, c_int64]),
    "xsmm_binary_invoke": (None, [c_int64, c_int64] + [c_void_p, c_int64] * 3),
    "xsmm_brgemm_invoke": (None, [c_int64, c_int64] + [c_void_p, c_int64] * 3 + [c_int64]),
    "xsmm_fused_brgemm_invoke": (None, [c_int64, c_int64] + [c_void_p, c_int64] * 4 + [c_int64]),
    "xsmm_intel_amx_tile_config_invoke": (None, [c_int64, c_int64, c_void_p, c_int64]),
    "perf_start_timer": (c_int64, []),
    "perf_stop_timer": (c_double, [c_int64]),
    "libxsmm_cpuid_dot_pack_factor": (c_int, [c_int]),
    "xsmm_cuda_set_stream": (None, [c_void_p]),
    "xsmm_cuda_get_stream": (c_void_p, []),
    "xsmm_cuda_stream_create": (c_void_p, []),
    "xsmm_cuda_stream_destroy": (None, [c_void_p]),
    "xsmm_cuda_sync": (None, []),
    "xsmm_cuda_stream_sync": (None, []),
    "xsmm_cuda_register_host": (c_int64, [c_void_p, c_int64, c_int64]),
    "xsmm_cuda_unregister_host": (c_int64, [c_void_p]),
    "xsmm_cuda_update_device": (c_int64, [c_void_p, c_int64]),
    "xsmm_cuda_update_host": (c_int64, [c_void_p, c_int64]),
    "xsmm_cuda_upload_async": (c_int64, [c_void_p, c_int64]),
    "xsmm_cuda_download_async": (c_int64, [c_void_p, c_int64]),
    "xsmm_cuda_wait_host": (c_int64, [c_void_p]),
    "xsmm_cuda_device_ptr": (c_void_p, [c_void_p]),
    "xsmm_cuda_graph_begin": (c_int64, []),
    "xsmm_cuda_graph_end": (c_int64, []),
    "xsmm_cuda_graph_launch": (None, [c_int64]),
    "xsmm_cuda_graph_destroy": (None, [c_int64]),
    "xsmm_cuda_set_lazy": (None, [c_int64]),
    "xsmm_cuda_mark_temporary": (None, [c_void_p, c_int64]),
    "xsmm_cuda_unmark_temporary": (None, [c_void_p]),
    "xsmm_cuda_debug_fold_grid": (c_int64, [c_int64] * 11 + [c_void_p] * 5),
    "xsmm_cuda_launch_count": (c_int64, []),
    "xsmm_cuda_last_kernel": (c_char_p, []),
    "xsmm_cuda_handle_kernel": (c_char_p, [c_int64]),
    "xsmm_cuda_abi_version": (c_int64, []),
    "xsmm_cuda_debug_dump_trace": (None, []),
    "xsmm_cuda_debug_tile_grid": (c_int64, [c_int64, c_void_p, c_void_p, c_void_p]),
    "xsmm_cuda_debug_rects_overlap": (c_int64, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64]),
}


def _load() -> ctypes.CDLL:
    path = os.environ.get("TPP_XSMM_LIB") or _build.lib_path()
    if not os.path.exists(path):
        # build in-tree if a toolchain is here; otherwise fail loudly (no CPU path exists)
        path = _build.build()
    try:
        lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    except OSError as e:  # pragma: no cover - depends on the box
        raise ImportError(f"tpp_mlir_b200: cannot load the CUDA backend {path}: {e}. "
                          "There is no CPU fallback; build it with `python -m tpp_mlir_b200._build`.") from e
    for name, (restype, argtypes) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if an ABI symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


LIB = _load()
LIB_PATH = LIB._name


def _ptr(x) -> int | None:
    """Address of the first element of a tensor-like (the memref's aligned pointer)."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):  # torch.Tensor
        return x.data_ptr()
    if hasattr(x, "ctypes"):  # numpy
        return x.ctypes.data
    raise TypeError(f"cannot take the address of {type(x)}")


def _sync_stream_with_torch(*tensors) -> None:
    """Launch on the stream torch would use for these tensors (device pointers only)."""
    for t in tensors:
        if hasattr(t, "is_cuda") and t.is_cuda:
            import torch

            LIB.xsmm_cuda_set_stream(torch.cuda.current_stream(t.device).cuda_stream)
            return


# ---- dispatch ---------------------------------------------------------------------
def gemm_dispatch(dtype, m, n, k, lda, ldb, ldc, flags=0) -> int:
    return LIB.xsmm_gemm_dispatch(dtype, m, n, k, lda, ldb, ldc, flags)


def brgemm_dispatch(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, flags=0) -> int:
    return LIB.xsmm_brgemm_dispatch(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, flags)


def fused_brgemm_dispatch(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, gemm_flags, unary_flags, unary_kind,
                          binary_flags, binary_kind) -> int:
    return LIB.xsmm_fused_brgemm_dispatch(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, gemm_flags, unary_flags,
                                          unary_kind, binary_flags, binary_kind)


def unary_dispatch(kind, dtype, m, n, ldi, ldo, flags=0) -> int:
    return LIB.xsmm_unary_dispatch(kind, dtype, m, n, ldi, ldo, flags)


def binary_dispatch(kind, dtype, m, n, ldi_lhs, ldi_rhs, ldo, flags=0) -> int:
    return LIB.xsmm_binary_dispatch(kind, dtype, m, n, ldi_lhs, ldi_rhs, ldo, flags)


def intel_amx_tile_config_dispatch(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, flags=0) -> int:
    return LIB.xsmm_intel_amx_tile_config_dispatch(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, flags)


# ---- invoke -----------------------------------------------------------------------
def gemm_invoke(dtype, handle, A, off_a, B, off_b, C, off_c) -> None:
    _sync_stream_with_torch(C, A, B)
    LIB.xsmm_gemm_invoke(dtype, handle, _ptr(A), off_a, _ptr(B), off_b, _ptr(C), off_c)


def brgemm_invoke(dtype, handle, A, off_a, B, off_b, C, off_c, num_batches) -> None:
    _sync_stream_with_torch(C, A, B)
    LIB.xsmm_brgemm_invoke(dtype, handle, _ptr(A), off_a, _ptr(B), off_b, _ptr(C), off_c, num_batches)


def fused_brgemm_invoke(dtype, handle, A, off_a, B, off_b, C, off_c, D, off_d, num_batches) -> None:
    _sync_stream_with_torch(C, A, B, D)
    LIB.xsmm_fused_brgemm_invoke(dtype, handle, _ptr(A), off_a, _ptr(B), off_b, _ptr(C), off_c, _ptr(D), off_d,
                                 num_batches)


def unary_invoke(dtype, handle, inp, off_in, out, off_out) -> None:
    _sync_stream_with_torch(out, inp)
    LIB.xsmm_unary_invoke(dtype, handle, _ptr(inp), off_in, _ptr(out), off_out)


def unary_scalar_invoke(dtype, handle, scalar, out, off_out) -> None:
    _sync_stream_with_torch(out)
    LIB.xsmm_unary_scalar_invoke(dtype, handle, float(scalar), _ptr(out), off_out)


def binary_invoke(dtype, handle, lhs, off_l, rhs, off_r, out, off_out) -> None:
    _sync_stream_with_torch(out, lhs, rhs)
    LIB.xsmm_binary_invoke(dtype, handle, _ptr(lhs), off_l, _ptr(rhs), off_r, _ptr(out), off_out)


def intel_amx_tile_config_invoke(dtype, handle, state, offset) -> None:
    LIB.xsmm_intel_amx_tile_config_invoke(dtype, handle, _ptr(state), offset)


# ---- timers / extensions ------------------------------------------------------------
def perf_start_timer() -> int:
    return LIB.perf_start_timer()


def perf_stop_timer(t0: int) -> float:
    return LIB.perf_stop_timer(t0)


def sync() -> None:
    LIB.xsmm_cuda_sync()


def set_lazy(on: bool) -> None:
    """Lazy mode of the calling thread: invokes on device operands are queued and launched fused at the next flush
    point (sync(), perf timers, update_* ...). Turning it off flushes."""
    LIB.xsmm_cuda_set_lazy(1 if on else 0)


def mark_temporary(t) -> None:
    """The tensor's memory is a function-local temporary of the program being replayed (see the ABI header): fused
    launches may drop its contents after their last use instead of writing them back to HBM."""
    nbytes = t.numel() * t.element_size() if hasattr(t, "element_size") else t.nbytes
    LIB.xsmm_cuda_mark_temporary(_ptr(t), nbytes)


def unmark_temporary(t) -> None:
    LIB.xsmm_cuda_unmark_temporary(_ptr(t))


def set_stream(stream_ptr: int | None) -> None:
    LIB.xsmm_cuda_set_stream(stream_ptr)


def get_stream() -> int | None:
    return LIB.xsmm_cuda_get_stream()


def stream_create() -> int:
    """A new non-blocking CUDA stream (as an integer handle) for set_stream; see xsmm_cuda_stream_create."""
    return LIB.xsmm_cuda_stream_create()


def stream_destroy(stream_ptr: int) -> None:
    LIB.xsmm_cuda_stream_destroy(stream_ptr)


def stream_sync() -> None:
    LIB.xsmm_cuda_stream_sync()


class graph_capture:
    """``with xsmm.graph_capture() as g: ...invokes...`` then ``g.launch()``: the captured invoke
    sequence (e.g. the body of a perf.bench loop) replays with one host call."""

    def __init__(self):
        self.handle = 0

    def __enter__(self):
        if LIB.xsmm_cuda_graph_begin() != 0:
            raise RuntimeError("xsmm_cuda_graph_begin failed (already capturing?)")
        return self

    def __exit__(self, exc_type, exc, tb):
        self.handle = LIB.xsmm_cuda_graph_end()
        if exc_type is None and not self.handle:
            raise RuntimeError("CUDA graph capture of the invoke sequence failed")
        return False

    def launch(self) -> None:
        LIB.xsmm_cuda_graph_launch(self.handle)

    def destroy(self) -> None:
        if self.handle:
            LIB.xsmm_cuda_graph_destroy(self.handle)
            self.handle = 0


def launch_count() -> int:
    return LIB.xsmm_cuda_launch_count()


def last_kernel() -> str:
    return LIB.xsmm_cuda_last_kernel().decode()


def handle_kernel(handle: int) -> str:
    return LIB.xsmm_cuda_handle_kernel(handle).decode()


def register_host(t, upload: bool = True) -> None:
    """Pin + mirror a CPU tensor/array on the device (weights, activations buffers)."""
    nbytes = t.numel() * t.element_size() if hasattr(t, "numel") else t.nbytes
    if LIB.xsmm_cuda_register_host(_ptr(t), nbytes, 1 if upload else 0) != 0:
        raise RuntimeError("xsmm_cuda_register_host failed")


def unregister_host(t) -> None:
    LIB.xsmm_cuda_unregister_host(_ptr(t))


def update_device(t, nbytes: int | None = None) -> None:
    n = nbytes if nbytes is not None else (t.numel() * t.element_size() if hasattr(t, "numel") else t.nbytes)
    if LIB.xsmm_cuda_update_device(_ptr(t), n) != 0:
        raise RuntimeError("xsmm_cuda_update_device: range is not registered")


def update_host(t, nbytes: int | None = None) -> None:
    n = nbytes if nbytes is not None else (t.numel() * t.element_size() if hasattr(t, "numel") else t.nbytes)
    if LIB.xsmm_cuda_update_host(_ptr(t), n) != 0:
        raise RuntimeError("xsmm_cuda_update_host: range is not registered")
