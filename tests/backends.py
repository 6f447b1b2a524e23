"""Two executors with one interface, so every parity case is written once.

* ``OracleBackend``  - the CPU restatement under oracle/ (the checker).
* ``AbiBackend``     - the product: calls go through the C-ABI of
  libtpp_xsmm_runner_utils.so exactly as lowered tpp-mlir code would call it.
  ``placement`` selects how operands are handed over:
    "device" - device pointers (fast path),
    "host"   - plain host pointers (strict mode: staged copies, synchronous),
    "mirror" - host pointers registered with xsmm_cuda_register_host.

Operands are flat numpy arrays (float32, or uint16 holding bf16 bits) plus an
element offset, like a memref's (aligned pointer, offset) pair. Outputs are
updated in place.
"""
from __future__ import annotations

import numpy as np

import oracle

F32, BF16 = 1, 2


def np_dtype(dtype):
    return np.float32 if dtype == F32 else np.uint16


class OracleBackend:
    name = "oracle"

    @staticmethod
    def _v(a, off):
        return None if a is None else a.reshape(-1)[off:]

    def brgemm(self, dtype, m, n, k, lda, ldb, ldc, sa, sb, flags, A, offA, B, offB, C, offC, batch):
        oracle.brgemm(dtype, m, n, k, lda, ldb, ldc, sa, sb, flags, self._v(A, offA), self._v(B, offB),
                      self._v(C, offC), batch)

    def gemm(self, dtype, m, n, k, lda, ldb, ldc, flags, A, offA, B, offB, C, offC):
        oracle.gemm(dtype, m, n, k, lda, ldb, ldc, flags, self._v(A, offA), self._v(B, offB), self._v(C, offC))

    def fused_brgemm(self, dtype, m, n, k, lda, ldb, ldc, sa, sb, gflags, uflags, ukind, bflags, bkind, A, offA, B,
                     offB, C, offC, D, offD, batch):
        oracle.fused_brgemm(dtype, m, n, k, lda, ldb, ldc, sa, sb, gflags, uflags, ukind, bflags, bkind,
                            self._v(A, offA), self._v(B, offB), self._v(C, offC), self._v(D, offD), batch)

    def unary(self, kind, dtype, m, n, ldi, ldo, flags, inp, offI, out, offO):
        src = self._v(inp, offI)
        if src is not None and src is not inp and np.shares_memory(src, out):
            src = src.copy()  # the oracle loops are not alias-safe for transposes; elementwise is
        oracle.unary(kind, dtype, m, n, ldi, ldo, flags, src, self._v(out, offO))

    def unary_scalar(self, kind, dtype, m, n, ldi, ldo, flags, scalar, out, offO):
        # by-value scalar rounded once to the element type, then the bcast_scalar kernel
        s = np.array([scalar], dtype=np.float32)
        if dtype == BF16:
            s = oracle.f32_to_bf16(s)
        oracle.unary(kind, dtype, m, n, 1, ldo, 8, s, self._v(out, offO))

    def binary(self, kind, dtype, m, n, ldl, ldr, ldo, flags, lhs, offL, rhs, offR, out, offO):
        oracle.binary(kind, dtype, m, n, ldl, ldr, ldo, flags, self._v(lhs, offL), self._v(rhs, offR),
                      self._v(out, offO))


class AbiBackend:
    """Calls the product through its C-ABI. Requires a GPU."""

    def __init__(self, placement="device"):
        import torch

        from tpp_mlir_b200 import xsmm

        self.torch = torch
        self.x = xsmm
        self.placement = placement
        self.name = f"abi-{placement}"
        self.kernels = []  # kernel variant names launched, in order

    # -- operand hand-over --------------------------------------------------------
    def _give(self, arrays):
        """arrays: list of numpy arrays (or None). Returns (handles, finalize)."""
        torch = self.torch
        handles, cleanup = [], []
        seen = {}
        for a in arrays:
            if a is None:
                handles.append(None)
                continue
            key = a.__array_interface__["data"][0]
            if key in seen:  # aliased operands (in-place ops) must alias on the device too
                handles.append(seen[key])
                continue
            flat = a.reshape(-1)
            assert flat.base is a or flat is a or np.shares_memory(flat, a)
            if self.placement == "device":
                t = torch.from_numpy(flat.view(np.int16) if a.dtype == np.uint16 else flat).cuda()
                cleanup.append(("d2h", t, flat))
                h = t
            elif self.placement == "host":
                h = flat  # plain pageable host memory
            else:  # mirror
                self.x.register_host(flat, upload=True)
                cleanup.append(("mirror", flat, None))
                h = flat
            seen[key] = h
            handles.append(h)

        def finalize():
            self.x.sync()
            for kind, obj, dst in cleanup:
                if kind == "d2h":
                    back = obj.cpu().numpy()
                    dst[:] = back.view(np.uint16) if dst.dtype == np.uint16 else back
                else:
                    self.x.update_host(obj)
                    self.x.sync()
                    self.x.unregister_host(obj)

        return handles, finalize

    def _ran(self):
        self.kernels.append(self.x.last_kernel())

    # -- ops ---------------------------------------------------------------------
    def brgemm(self, dtype, m, n, k, lda, ldb, ldc, sa, sb, flags, A, offA, B, offB, C, offC, batch):
        h = self.x.brgemm_dispatch(dtype, m, n, k, lda, ldb, ldc, sa, sb, flags)
        (a, b, c), fin = self._give([A, B, C])
        self.x.brgemm_invoke(dtype, h, a, offA, b, offB, c, offC, batch)
        self._ran()
        fin()

    def gemm(self, dtype, m, n, k, lda, ldb, ldc, flags, A, offA, B, offB, C, offC):
        h = self.x.gemm_dispatch(dtype, m, n, k, lda, ldb, ldc, flags)
        (a, b, c), fin = self._give([A, B, C])
        self.x.gemm_invoke(dtype, h, a, offA, b, offB, c, offC)
        self._ran()
        fin()

    def fused_brgemm(self, dtype, m, n, k, lda, ldb, ldc, sa, sb, gflags, uflags, ukind, bflags, bkind, A, offA, B,
                     offB, C, offC, D, offD, batch):
        h = self.x.fused_brgemm_dispatch(dtype, m, n, k, lda, ldb, ldc, sa, sb, gflags, uflags, ukind, bflags, bkind)
        (a, b, c, d), fin = self._give([A, B, C, D])
        self.x.fused_brgemm_invoke(dtype, h, a, offA, b, offB, c, offC, d, offD, batch)
        self._ran()
        fin()

    def unary(self, kind, dtype, m, n, ldi, ldo, flags, inp, offI, out, offO):
        h = self.x.unary_dispatch(kind, dtype, m, n, ldi, ldo, flags)
        (i, o), fin = self._give([inp, out])
        self.x.unary_invoke(dtype, h, i, offI, o, offO)
        self._ran()
        fin()

    def unary_scalar(self, kind, dtype, m, n, ldi, ldo, flags, scalar, out, offO):
        h = self.x.unary_dispatch(kind, dtype, m, n, ldi, ldo, flags)
        (o,), fin = self._give([out])
        self.x.unary_scalar_invoke(dtype, h, scalar, o, offO)
        self._ran()
        fin()

    def binary(self, kind, dtype, m, n, ldl, ldr, ldo, flags, lhs, offL, rhs, offR, out, offO):
        h = self.x.binary_dispatch(kind, dtype, m, n, ldl, ldr, ldo, flags)
        (l, r, o), fin = self._give([lhs, rhs, out])
        self.x.binary_invoke(dtype, h, l, offL, r, offR, o, offO)
        self._ran()
        fin()
