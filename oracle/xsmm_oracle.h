/*
 * xsmm_oracle.h - CPU restatement of the xsmm TPP semantics (TEST INFRASTRUCTURE).
 *
 * This directory is the parity oracle for the B200 backend. It is test
 * infrastructure only: nothing under tpp_mlir_b200/ may include, link or call
 * it. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, and only as the checker / the CPU arm.
 *
 * The arithmetic of the reference lives in libxsmm @ 85851d43 (third party,
 * fetched by cmake/modules/xsmm.cmake:14-18, absent from /root/reference and
 * not buildable offline), so this file restates the published operator
 * semantics as seen through the reference's own call sites:
 *   runtime/Xsmm/XsmmRunnerUtils.cpp:79-93,95-140   gemm
 *   runtime/Xsmm/XsmmRunnerUtils.cpp:288-361        brgemm (stride variant)
 *   runtime/Xsmm/XsmmRunnerUtils.cpp:363-457        fused brgemm (+binary +unary)
 *   runtime/Xsmm/XsmmRunnerUtils.cpp:142-179,248-286 unary
 *   runtime/Xsmm/XsmmRunnerUtils.cpp:181-211,261-274 binary
 *   lib/TPP/Dialect/Xsmm/XsmmUtils.cpp:90-252       ld / broadcast rules
 *   lib/TPP/Transforms/Utils/VNNIUtils.cpp:75-78    VNNI layouts
 * and is pinned against the known-answer vectors of test/Integration/xsmm-*.mlir
 * and the .mlir files under test/BF16/Integration (tests/test_oracle_golden.py).
 *
 * Everything is the ROW-MAJOR view that tpp-mlir passes (the shim's swap to
 * libxsmm's column-major view is an implementation detail of the reference).
 */
#ifndef XSMM_ORACLE_H
#define XSMM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* bf16 <-> f32, round-to-nearest-even (mlir Float16bits.h semantics; the
 * reference's storage type for bf16 operands, XsmmRunnerUtils.cpp:68). */
uint16_t xo_f32_to_bf16(float f);
float xo_bf16_to_f32(uint16_t h);
void xo_f32_to_bf16_array(const float *src, uint16_t *dst, int64_t n);
void xo_bf16_to_f32_array(const uint16_t *src, float *dst, int64_t n);

/* acc_mode: 0 = f32 accumulation (the reference's comp_type,
 * XsmmRunnerUtils.cpp:342-343), 1 = f64 accumulation (tighter truth used by the
 * f32 tolerance tests; summation order is unspecified in the reference). */
void xo_set_acc_mode(int mode);
/* number of OpenMP threads the oracle will use (1 when built without OpenMP) */
int xo_num_threads(void);
/* VNNI blocking factor of B operands with gemm flag 2048 (2 by default, 4 for mlir-gen --vnni=4 layouts) */
void xo_set_vnni_factor(int v);
int xo_vnni_factor(void);
void xo_set_num_threads(int n);

/* C (+)= sum_b A_b * B_b ; flags as received by the C-ABI (see tpp_xsmm_abi.h) */
void xo_brgemm(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda,
               int64_t ldb, int64_t ldc, int64_t stride_a, int64_t stride_b,
               int64_t flags, const void *A, const void *B, void *C,
               int64_t batch);

void xo_gemm(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda,
             int64_t ldb, int64_t ldc, int64_t flags, const void *A,
             const void *B, void *C);

/* C = unary( binary( [C +] sum_b A_b*B_b , D ) ), post-ops on the f32
 * accumulator, ONE rounding at the store (XsmmRunnerUtils.cpp:430-446). */
void xo_fused_brgemm(int64_t dtype, int64_t m, int64_t n, int64_t k,
                     int64_t lda, int64_t ldb, int64_t ldc, int64_t stride_a,
                     int64_t stride_b, int64_t gemm_flags, int64_t unary_flags,
                     int64_t unary_kind, int64_t binary_flags,
                     int64_t binary_kind, const void *A, const void *B, void *C,
                     const void *D, int64_t batch);

/* returns 0 on success, -1 for an unsupported kind/flag combination */
int xo_unary(int64_t kind, int64_t dtype, int64_t m, int64_t n, int64_t ldi,
             int64_t ldo, int64_t flags, const void *in, void *out);

int xo_binary(int64_t kind, int64_t dtype, int64_t m, int64_t n, int64_t ldl,
              int64_t ldr, int64_t ldo, int64_t flags, const void *lhs,
              const void *rhs, void *out);

/* xsmm_oracle_fast.c: vectorised / AMX variants for the timed CPU arm; 0 on success, -1 if unsupported here */
int xo_fused_brgemm_fast(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc,
                         int64_t stride_a, int64_t stride_b, int64_t gemm_flags, int64_t unary_kind,
                         int64_t binary_flags, int64_t binary_kind, const void *A, const void *B, void *C,
                         const void *D, int64_t batch);
int xo_fused_brgemm_amx(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc,
                        int64_t stride_a, int64_t stride_b, int64_t gemm_flags, int64_t unary_kind,
                        int64_t binary_flags, int64_t binary_kind, const void *A, const void *B, void *C,
                        const void *D, int64_t batch);
int xo_fused_brgemm_amx_grid(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc,
                             int64_t stride_a, int64_t stride_b, int64_t gemm_flags, int64_t unary_kind,
                             int64_t binary_flags, int64_t binary_kind, const void *A, const void *B, void *C, const void *D,
                             int64_t batch, int64_t grid_n, int64_t grid_k, int64_t a_step, int64_t b_step, int64_t c_step_n,
                             int64_t c_step_k, int64_t d_step);
int xo_fast_isa(void);

#ifdef __cplusplus
}
#endif
#endif
