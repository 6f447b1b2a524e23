/*
 * xsmm_oracle_fast.c - a vectorised CPU implementation of the bf16 fused BRGEMM, used ONLY as the timed CPU arm
 * of bench.py (cpu_baseline / --impl reference). TEST INFRASTRUCTURE, never linked by the product.
 *
 * The reference's CPU path is libxsmm's JIT (AVX512-BF16 / AMX microkernels, runtime/Xsmm/XsmmRunnerUtils.cpp:
 * 385-457); libxsmm cannot be built offline, and the plain-C oracle (xsmm_oracle.c) is written for clarity, not
 * speed, so timing it would flatter the GPU. This file is the closest honest stand-in: the same operator
 * (C = relu(beta*C + sum_b A_b*B_b + bias), f32 accumulation, one RNE rounding) with an 8 x 32 register-blocked
 * microkernel on vdpbf16ps (AVX512-BF16) when the build machine has it, else on AVX-512/AVX2 f32 FMAs through
 * the compiler's vectoriser. OpenMP over output tiles. It is validated against the plain oracle in
 * tests/test_oracle_golden.py::test_fast_cpu_kernel_matches_oracle (1e-2 rel; vdpbf16ps sums pairs of products
 * before adding to the accumulator, so the last bits differ from the scalar order).
 *
 * Supported: dtype bf16, flat (non-VNNI) B, beta_0 or accumulate, optional bias (bcast_col_in0 add) and relu.
 * Returns -1 for anything else (the caller then uses the plain oracle).
 */
#include "xsmm_oracle.h"

#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__AVX512F__) && defined(__AVX512BF16__) && defined(__AVX512BW__)
#include <immintrin.h>
#define XO_FAST_BF16 1
#else
#define XO_FAST_BF16 0
#endif

int xo_fast_isa(void) { return XO_FAST_BF16 ? 2 : 1; } /* 2 = AVX512-BF16 microkernel, 1 = compiler-vectorised f32 */

#define TM 8
#define TN 32

#if XO_FAST_BF16
/* one 8 x 32 output tile, all batches and k (k even) */
static void tile_bf16(const uint16_t *A, const uint16_t *B, int64_t k, int64_t lda, int64_t ldb, int64_t stride_a,
                      int64_t stride_b, int64_t batch, float acc[TM][TN]) {
  __m512 c[TM][2];
  for (int r = 0; r < TM; ++r) {
    c[r][0] = _mm512_loadu_ps(&acc[r][0]);
    c[r][1] = _mm512_loadu_ps(&acc[r][16]);
  }
  /* interleave two rows of 32 bf16 into pairs (b[p][j], b[p+1][j]) for j = 0..15 and 16..31 */
  const __m512i idx_lo = _mm512_set_epi16(47, 15, 46, 14, 45, 13, 44, 12, 43, 11, 42, 10, 41, 9, 40, 8, 39, 7, 38, 6, 37,
                                          5, 36, 4, 35, 3, 34, 2, 33, 1, 32, 0);
  const __m512i idx_hi = _mm512_set_epi16(63, 31, 62, 30, 61, 29, 60, 28, 59, 27, 58, 26, 57, 25, 56, 24, 55, 23, 54, 22,
                                          53, 21, 52, 20, 51, 19, 50, 18, 49, 17, 48, 16);
  for (int64_t b = 0; b < batch; ++b) {
    const uint16_t *Ab = A + b * stride_a, *Bb = B + b * stride_b;
    for (int64_t p = 0; p < k; p += 2) {
      const __m512i r0 = _mm512_loadu_si512((const void *)(Bb + p * ldb));       /* 32 bf16 of row p   */
      const __m512i r1 = _mm512_loadu_si512((const void *)(Bb + (p + 1) * ldb)); /* 32 bf16 of row p+1 */
      const __m512bh b0 = (__m512bh)_mm512_permutex2var_epi16(r0, idx_lo, r1);
      const __m512bh b1 = (__m512bh)_mm512_permutex2var_epi16(r0, idx_hi, r1);
      for (int r = 0; r < TM; ++r) {
        int32_t pair;
        memcpy(&pair, Ab + r * lda + p, 4);
        const __m512bh a = (__m512bh)_mm512_set1_epi32(pair);
        c[r][0] = _mm512_dpbf16_ps(c[r][0], a, b0);
        c[r][1] = _mm512_dpbf16_ps(c[r][1], a, b1);
      }
    }
  }
  for (int r = 0; r < TM; ++r) {
    _mm512_storeu_ps(&acc[r][0], c[r][0]);
    _mm512_storeu_ps(&acc[r][16], c[r][1]);
  }
}
#else
static void tile_bf16(const uint16_t *A, const uint16_t *B, int64_t k, int64_t lda, int64_t ldb, int64_t stride_a,
                      int64_t stride_b, int64_t batch, float acc[TM][TN]) {
  for (int64_t b = 0; b < batch; ++b) {
    const uint16_t *Ab = A + b * stride_a, *Bb = B + b * stride_b;
    for (int64_t p = 0; p < k; ++p) {
      float brow[TN];
      for (int j = 0; j < TN; ++j) brow[j] = xo_bf16_to_f32(Bb[p * ldb + j]);
      for (int r = 0; r < TM; ++r) {
        const float a = xo_bf16_to_f32(Ab[r * lda + p]);
        for (int j = 0; j < TN; ++j) acc[r][j] += a * brow[j];
      }
    }
  }
}
#endif

int xo_fused_brgemm_fast(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc,
                         int64_t stride_a, int64_t stride_b, int64_t gemm_flags, int64_t unary_kind,
                         int64_t binary_flags, int64_t binary_kind, const void *A, const void *B, void *C,
                         const void *D, int64_t batch) {
  if (dtype != 2 || (gemm_flags & 2048) || (m % TM) || (n % TN) || (k % 2)) return -1;
  if (binary_kind != 0 && !(binary_kind == 1 && binary_flags == 4)) return -1;
  if (unary_kind != 0 && unary_kind != 5) return -1;
  const int beta0 = (gemm_flags & 4) != 0;
  const uint16_t *Ap = (const uint16_t *)A, *Bp = (const uint16_t *)B, *Dp = (const uint16_t *)D;
  uint16_t *Cp = (uint16_t *)C;
  const int64_t tm = m / TM, tn = n / TN;
#ifdef _OPENMP
#pragma omp parallel for collapse(2) schedule(static) num_threads(xo_num_threads())
#endif
  for (int64_t bi = 0; bi < tm; ++bi) {
    for (int64_t bj = 0; bj < tn; ++bj) {
      float acc[TM][TN];
      for (int r = 0; r < TM; ++r)
        for (int j = 0; j < TN; ++j)
          acc[r][j] = beta0 ? 0.0f : xo_bf16_to_f32(Cp[(bi * TM + r) * ldc + bj * TN + j]);
      tile_bf16(Ap + bi * TM * lda, Bp + bj * TN, k, lda, ldb, stride_a, stride_b, batch, acc);
      for (int r = 0; r < TM; ++r)
        for (int j = 0; j < TN; ++j) {
          float v = acc[r][j];
          if (binary_kind == 1 && Dp) v += xo_bf16_to_f32(Dp[bj * TN + j]);
          if (unary_kind == 5) v = v > 0.0f ? v : 0.0f;
          Cp[(bi * TM + r) * ldc + bj * TN + j] = xo_f32_to_bf16(v);
        }
    }
  }
  return 0;
}
