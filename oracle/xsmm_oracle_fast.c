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
 *
 * xo_fused_brgemm_amx: the same operator on AMX-BF16 tiles (tdpbf16ps, 2 x 2 blocking of 16 x 16 f32 tiles, K = 32
 * per step) for a VNNI-2 packed B ([k/2][n][2], gemm flag 2048) - the layout and the instruction libxsmm's JIT uses
 * for bf16 on Sapphire Rapids and later (the reference packs its weights with --vnni=2 for exactly this reason,
 * benchmarks/config/omp/mlir-bf16.json:37). Built only when the compiler targets AMX (-march=native on such a host);
 * needs the kernel's permission for tile state (arch_prctl ARCH_REQ_XCOMP_PERM). Returns -1 when unavailable.
 */
#define _GNU_SOURCE /* syscall() for the AMX permission request */
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

#if defined(__AMX_TILE__) && defined(__AMX_BF16__) && defined(__AVX512F__) && defined(__AVX512BF16__) && defined(__linux__)
#include <immintrin.h>
#include <sys/syscall.h>
#include <unistd.h>
#define XO_FAST_AMX 1
#else
#define XO_FAST_AMX 0
#endif

#if XO_FAST_AMX
static int amx_state = 0; /* 0 = not asked yet, 1 = usable, -1 = refused */
static int amx_usable(void) {
  if (amx_state == 0) {
    /* ARCH_REQ_XCOMP_PERM = 0x1023, XFEATURE_XTILEDATA = 18 */
    amx_state = syscall(SYS_arch_prctl, 0x1023, 18) == 0 ? 1 : -1;
  }
  return amx_state == 1;
}
#else
static int amx_usable(void) { return 0; }
#endif

/* 3 = AMX-BF16 tiles available, 2 = AVX512-BF16 microkernel, 1 = compiler-vectorised f32 */
int xo_fast_isa(void) { return amx_usable() ? 3 : XO_FAST_BF16 ? 2 : 1; }

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

#if XO_FAST_AMX
typedef struct {
  uint8_t palette_id, start_row, reserved[14];
  uint16_t colsb[16];
  uint8_t rows[16];
} __attribute__((packed, aligned(64))) xo_tilecfg;

/* one 32 x 32 output block: C tiles 0..3 (2 x 2), A tiles 4, 5 (16 rows x 32 k), B tiles 6, 7 (16 k pairs x 16 columns x 2) */
static void amx_block(const uint16_t *A, const uint16_t *B, int64_t k, int64_t lda, int64_t ldb, int64_t stride_a,
                      int64_t stride_b, int64_t batch, float *acc /* [32][32] */) {
  _tile_zero(0); _tile_zero(1); _tile_zero(2); _tile_zero(3);
  for (int64_t b = 0; b < batch; ++b) {
    const uint16_t *Ab = A + b * stride_a, *Bb = B + b * stride_b;
    for (int64_t p = 0; p < k; p += 32) {
      _tile_loadd(4, Ab + p, lda * 2);
      _tile_loadd(5, Ab + 16 * lda + p, lda * 2);
      _tile_loadd(6, Bb + (p / 2) * ldb * 2, ldb * 4);
      _tile_loadd(7, Bb + (p / 2) * ldb * 2 + 32, ldb * 4);
      _tile_dpbf16ps(0, 4, 6);
      _tile_dpbf16ps(1, 4, 7);
      _tile_dpbf16ps(2, 5, 6);
      _tile_dpbf16ps(3, 5, 7);
    }
  }
  _tile_stored(0, acc, 32 * 4);
  _tile_stored(1, acc + 16, 32 * 4);
  _tile_stored(2, acc + 16 * 32, 32 * 4);
  _tile_stored(3, acc + 16 * 32 + 16, 32 * 4);
}
#endif

#if XO_FAST_AMX
static void amx_config(void) {
  xo_tilecfg cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.palette_id = 1;
  for (int t = 0; t < 8; ++t) { cfg.rows[t] = 16; cfg.colsb[t] = 64; }
  _tile_loadconfig(&cfg);
}

/* all 32 x 32 blocks of one m x n BRGEMM: accumulate on the tiles, then (+C) + bias -> relu -> bf16 */
static void amx_tile_brgemm(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc, int64_t stride_a,
                            int64_t stride_b, int beta0, int relu, const uint16_t *Ap, const uint16_t *Bp, uint16_t *Cp,
                            const uint16_t *Dp, int64_t batch, float *acc) {
  for (int64_t bi = 0; bi < m / 32; ++bi) {
    for (int64_t bj = 0; bj < n / 32; ++bj) {
      /* B in VNNI-2: element (p, j) at ((p / 2) * ldb + j) * 2 + p % 2; a column block starts 2 * 32 bf16 further */
      amx_block(Ap + bi * 32 * lda, Bp + bj * 32 * 2, k, lda, ldb, stride_a, stride_b, batch, acc);
      for (int r = 0; r < 32; ++r) {
        uint16_t *crow = Cp + (bi * 32 + r) * ldc + bj * 32;
        for (int h = 0; h < 2; ++h) {
          __m512 v = _mm512_load_ps(acc + r * 32 + 16 * h);
          if (!beta0)
            v = _mm512_add_ps(v, _mm512_castsi512_ps(_mm512_slli_epi32(
                                     _mm512_cvtepu16_epi32(_mm256_loadu_si256((const __m256i *)(crow + 16 * h))), 16)));
          if (Dp)
            v = _mm512_add_ps(v, _mm512_castsi512_ps(_mm512_slli_epi32(
                                     _mm512_cvtepu16_epi32(_mm256_loadu_si256((const __m256i *)(Dp + bj * 32 + 16 * h))), 16)));
          if (relu) v = _mm512_max_ps(v, _mm512_setzero_ps());
          _mm256_storeu_si256((__m256i *)(crow + 16 * h), (__m256i)_mm512_cvtneps_pbh(v));
        }
      }
    }
  }
}
#endif

/* grid_n x grid_k tile BRGEMMs of ONE shape - tile (i, j) on A + i a_step, B + j b_step, C + i c_step_n + j c_step_k,
 * D + j d_step (elements) - distributed over the OpenMP threads: the reference's scf.parallel loop nest over the (iN, iK)
 * output blocks of a layer, each iteration one libxsmm BRGEMM call. 1 x 1 = a single invoke, whose
 * 32 x 32 blocks are then what the threads share. */
int xo_fused_brgemm_amx_grid(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc,
                             int64_t stride_a, int64_t stride_b, int64_t gemm_flags, int64_t unary_kind,
                             int64_t binary_flags, int64_t binary_kind, const void *A, const void *B, void *C, const void *D,
                             int64_t batch, int64_t grid_n, int64_t grid_k, int64_t a_step, int64_t b_step, int64_t c_step_n,
                             int64_t c_step_k, int64_t d_step) {
#if XO_FAST_AMX
  if (!amx_usable()) return -1;
  if (dtype != 2 || !(gemm_flags & 2048) || (m % 32) || (n % 32) || (k % 32)) return -1;
  if (binary_kind != 0 && !(binary_kind == 1 && binary_flags == 4)) return -1;
  if (unary_kind != 0 && unary_kind != 5) return -1;
  const int beta0 = (gemm_flags & 4) != 0, relu = unary_kind == 5;
  const uint16_t *Ap = (const uint16_t *)A, *Bp = (const uint16_t *)B, *Dp = binary_kind == 1 ? (const uint16_t *)D : NULL;
  uint16_t *Cp = (uint16_t *)C;
  if (grid_n == 1 && grid_k == 1) {
    /* one invoke: its 32 x 32 blocks are the parallel work */
    const int64_t tm = m / 32, tn = n / 32;
#ifdef _OPENMP
#pragma omp parallel num_threads(xo_num_threads())
#endif
    {
      amx_config();
      float acc[32 * 32] __attribute__((aligned(64)));
#ifdef _OPENMP
#pragma omp for collapse(2) schedule(static)
#endif
      for (int64_t bi = 0; bi < tm; ++bi)
        for (int64_t bj = 0; bj < tn; ++bj)
          amx_tile_brgemm(32, 32, k, lda, ldb, ldc, stride_a, stride_b, beta0, relu, Ap + bi * 32 * lda, Bp + bj * 64,
                          Cp + bi * 32 * ldc + bj * 32, Dp ? Dp + bj * 32 : NULL, batch, acc);
      _tile_release();
    }
    return 0;
  }
#ifdef _OPENMP
#pragma omp parallel num_threads(xo_num_threads())
#endif
  {
    amx_config();
    float acc[32 * 32] __attribute__((aligned(64)));
#ifdef _OPENMP
#pragma omp for collapse(2) schedule(static)
#endif
    for (int64_t i = 0; i < grid_n; ++i)
      for (int64_t j = 0; j < grid_k; ++j)
        amx_tile_brgemm(m, n, k, lda, ldb, ldc, stride_a, stride_b, beta0, relu, Ap + i * a_step, Bp + j * b_step,
                        Cp + i * c_step_n + j * c_step_k, Dp ? Dp + j * d_step : NULL, batch, acc);
    _tile_release();
  }
  return 0;
#else
  (void)dtype; (void)m; (void)n; (void)k; (void)lda; (void)ldb; (void)ldc; (void)stride_a; (void)stride_b; (void)gemm_flags;
  (void)unary_kind; (void)binary_flags; (void)binary_kind; (void)A; (void)B; (void)C; (void)D; (void)batch;
  (void)grid_n; (void)grid_k; (void)a_step; (void)b_step; (void)c_step_n; (void)c_step_k; (void)d_step;
  return -1;
#endif
}

int xo_fused_brgemm_amx(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc,
                        int64_t stride_a, int64_t stride_b, int64_t gemm_flags, int64_t unary_kind, int64_t binary_flags,
                        int64_t binary_kind, const void *A, const void *B, void *C, const void *D, int64_t batch) {
  return xo_fused_brgemm_amx_grid(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, gemm_flags, unary_kind, binary_flags,
                                  binary_kind, A, B, C, D, batch, 1, 1, 0, 0, 0, 0, 0);
}
