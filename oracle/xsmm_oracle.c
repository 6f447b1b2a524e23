/*
 * xsmm_oracle.c - CPU restatement of the xsmm TPPs. TEST INFRASTRUCTURE ONLY
 * (see xsmm_oracle.h for the provenance of every function and the rule that the
 * product path never links this).
 *
 * Plain C, f32 accumulation in a fixed (batch-major, then k) order per output
 * element, one RNE rounding at the bf16 store. OpenMP over row blocks when built
 * with -fopenmp (used for the cpu_baseline timing; results do not depend on the
 * thread count because each output element is owned by one thread and summed in
 * a fixed order).
 */
#include "xsmm_oracle.h"

#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define XO_F32 1
#define XO_BF16 2

#define XO_BETA_0 4
#define XO_B_VNNI 2048 /* row-major B operand is [K/v][N][v] */

/* VNNI blocking factor v of B operands: what libxsmm_cpuid_dot_pack_factor(BF16) answers on the target - 2 on every x86
 * part the reference's tests run on (and the default here), 4 where mlir-gen --vnni=4 applies
 * (lib/TPP/Transforms/Utils/VNNIUtils.cpp:25-45; benchmarks/config/omp/mlir-bf16.json:65-125). Process-wide, like the
 * cpuid answer it stands for. */
static int64_t g_vnni_factor = 2;
void xo_set_vnni_factor(int v) { g_vnni_factor = (v == 4) ? 4 : 2; }
int xo_vnni_factor(void) { return (int)g_vnni_factor; }

static int g_acc_mode = 0;
static int g_threads = 0;

void xo_set_acc_mode(int mode) { g_acc_mode = mode; }

int xo_num_threads(void) {
#ifdef _OPENMP
  return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
  return 1;
#endif
}

void xo_set_num_threads(int n) { g_threads = n; }

/* mlir/ExecutionEngine/Float16bits.h float2bfloat(): RNE on the upper 16 bits,
 * NaN kept quiet. */
uint16_t xo_f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) /* NaN */
    return (uint16_t)((u >> 16) | 0x0040u);
  uint32_t lsb = (u >> 16) & 1u;
  u += 0x7fffu + lsb;
  return (uint16_t)(u >> 16);
}

float xo_bf16_to_f32(uint16_t h) {
  uint32_t u = ((uint32_t)h) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

void xo_f32_to_bf16_array(const float *src, uint16_t *dst, int64_t n) {
  for (int64_t i = 0; i < n; ++i)
    dst[i] = xo_f32_to_bf16(src[i]);
}

void xo_bf16_to_f32_array(const uint16_t *src, float *dst, int64_t n) {
  for (int64_t i = 0; i < n; ++i)
    dst[i] = xo_bf16_to_f32(src[i]);
}

static inline float ld(int64_t dtype, const void *p, int64_t idx) {
  if (dtype == XO_F32)
    return ((const float *)p)[idx];
  return xo_bf16_to_f32(((const uint16_t *)p)[idx]);
}

static inline void st(int64_t dtype, void *p, int64_t idx, float v) {
  if (dtype == XO_F32)
    ((float *)p)[idx] = v;
  else
    ((uint16_t *)p)[idx] = xo_f32_to_bf16(v);
}

/* relu exactly as a select: negative, -0 and NaN all become +0 */
static inline float relu(float x) { return x > 0.0f ? x : 0.0f; }

/* One block of rows [i0,i1): acc[i][j] over all batches and k, f32. */
#define XO_ROWBLK 16
#define XO_COLBLK 128

static void brgemm_rows_f32acc(int64_t dtype, int64_t i0, int64_t i1, int64_t j0,
                               int64_t j1, int64_t k, int64_t lda, int64_t ldb,
                               int64_t ldc, int64_t stride_a, int64_t stride_b,
                               int vnni, int beta0, const void *A, const void *B,
                               const void *C, int64_t batch, float *acc,
                               float *brow) {
  const int64_t rows = i1 - i0, n = j1 - j0;
  for (int64_t r = 0; r < rows; ++r)
    for (int64_t j = 0; j < n; ++j)
      acc[r * n + j] = beta0 ? 0.0f : ld(dtype, C, (i0 + r) * ldc + j0 + j);
  for (int64_t b = 0; b < batch; ++b) {
    for (int64_t p = 0; p < k; ++p) {
      /* expand columns [j0,j1) of B row p of batch b into f32 once per tile */
      if (dtype == XO_F32) {
        const float *Bp = (const float *)B + b * stride_b + p * ldb + j0;
        for (int64_t j = 0; j < n; ++j)
          brow[j] = Bp[j];
      } else if (!vnni) {
        const uint16_t *Bp = (const uint16_t *)B + b * stride_b + p * ldb + j0;
        for (int64_t j = 0; j < n; ++j) {
          uint32_t u = ((uint32_t)Bp[j]) << 16;
          memcpy(&brow[j], &u, 4);
        }
      } else {
        /* B[b][p/v][j][p%v], ldb already divided by the VNNI factor v
         * (ConvertLinalgToXsmm.cpp:1143-1148) */
        const int64_t v = g_vnni_factor;
        const uint16_t *Bp = (const uint16_t *)B + b * stride_b +
                             ((p / v) * ldb + j0) * v + (p % v);
        for (int64_t j = 0; j < n; ++j) {
          uint32_t u = ((uint32_t)Bp[j * v]) << 16;
          memcpy(&brow[j], &u, 4);
        }
      }
      for (int64_t r = 0; r < rows; ++r) {
        const float a = ld(dtype, A, b * stride_a + (i0 + r) * lda + p);
        float *accr = acc + r * n;
        for (int64_t j = 0; j < n; ++j)
          accr[j] += a * brow[j];
      }
    }
  }
}

static double brgemm_elem_f64(int64_t dtype, int64_t i, int64_t j, int64_t k,
                              int64_t lda, int64_t ldb, int64_t stride_a,
                              int64_t stride_b, int vnni, const void *A,
                              const void *B, int64_t batch, double acc) {
  for (int64_t b = 0; b < batch; ++b)
    for (int64_t p = 0; p < k; ++p) {
      float a = ld(dtype, A, b * stride_a + i * lda + p);
      const int64_t v = g_vnni_factor;
      float bb = vnni ? ld(dtype, B, b * stride_b + ((p / v) * ldb + j) * v + p % v)
                      : ld(dtype, B, b * stride_b + p * ldb + j);
      acc += (double)a * (double)bb;
    }
  return acc;
}

void xo_fused_brgemm(int64_t dtype, int64_t m, int64_t n, int64_t k,
                     int64_t lda, int64_t ldb, int64_t ldc, int64_t stride_a,
                     int64_t stride_b, int64_t gemm_flags, int64_t unary_flags,
                     int64_t unary_kind, int64_t binary_flags,
                     int64_t binary_kind, const void *A, const void *B, void *C,
                     const void *D, int64_t batch) {
  (void)unary_flags;
  const int vnni = (gemm_flags & XO_B_VNNI) != 0 && dtype == XO_BF16;
  const int beta0 = (gemm_flags & XO_BETA_0) != 0;
  /* the only binary the lowering emits is ADD with bcast_col_in0: D is a
   * length-n vector (ConvertXsmmToFunc.cpp:405-422); other bcast modes of D are
   * restated for completeness: none -> D[i*ldc+j], row_in0 -> D[i], scalar. */
  const int has_bin = binary_kind != 0 && D != NULL;
  const int has_relu = unary_kind == 5;

  if (g_acc_mode == 1) {
    for (int64_t i = 0; i < m; ++i)
      for (int64_t j = 0; j < n; ++j) {
        double acc = beta0 ? 0.0 : (double)ld(dtype, C, i * ldc + j);
        acc = brgemm_elem_f64(dtype, i, j, k, lda, ldb, stride_a, stride_b, vnni,
                              A, B, batch, acc);
        float v = (float)acc;
        if (has_bin) {
          float d = (binary_flags & 4)    ? ld(dtype, D, j)
                    : (binary_flags & 1)  ? ld(dtype, D, i)
                    : (binary_flags & 16) ? ld(dtype, D, 0)
                                          : ld(dtype, D, i * ldc + j);
          v = binary_kind == 1   ? v + d
              : binary_kind == 2 ? v * d
              : binary_kind == 3 ? v - d
                                 : v / d;
        }
        if (has_relu)
          v = relu(v);
        st(dtype, C, i * ldc + j, v);
      }
    return;
  }

  /* tiles of XO_ROWBLK rows x XO_COLBLK columns; each output element is owned by
   * exactly one task and summed in a fixed order, so the thread count never
   * changes a result. */
  const int64_t nblk = (m + XO_ROWBLK - 1) / XO_ROWBLK;
  const int64_t ncb = (n + XO_COLBLK - 1) / XO_COLBLK;
#ifdef _OPENMP
  int nthr = xo_num_threads();
#pragma omp parallel num_threads(nthr)
#endif
  {
    float *acc = (float *)malloc(sizeof(float) * XO_ROWBLK * XO_COLBLK);
    float *brow = (float *)malloc(sizeof(float) * XO_COLBLK);
#ifdef _OPENMP
#pragma omp for schedule(static) collapse(2)
#endif
    for (int64_t blk = 0; blk < nblk; ++blk) {
      for (int64_t cb = 0; cb < ncb; ++cb) {
        const int64_t i0 = blk * XO_ROWBLK;
        const int64_t i1 = i0 + XO_ROWBLK < m ? i0 + XO_ROWBLK : m;
        const int64_t j0 = cb * XO_COLBLK;
        const int64_t j1 = j0 + XO_COLBLK < n ? j0 + XO_COLBLK : n;
        const int64_t w = j1 - j0;
        brgemm_rows_f32acc(dtype, i0, i1, j0, j1, k, lda, ldb, ldc, stride_a,
                           stride_b, vnni, beta0, A, B, C, batch, acc, brow);
        for (int64_t i = i0; i < i1; ++i)
          for (int64_t j = j0; j < j1; ++j) {
            float v = acc[(i - i0) * w + (j - j0)];
            if (has_bin) {
              float d = (binary_flags & 4)    ? ld(dtype, D, j)
                        : (binary_flags & 1)  ? ld(dtype, D, i)
                        : (binary_flags & 16) ? ld(dtype, D, 0)
                                              : ld(dtype, D, i * ldc + j);
              v = binary_kind == 1   ? v + d
                  : binary_kind == 2 ? v * d
                  : binary_kind == 3 ? v - d
                                     : v / d;
            }
            if (has_relu)
              v = relu(v);
            st(dtype, C, i * ldc + j, v);
          }
      }
    }
    free(acc);
    free(brow);
  }
}

void xo_brgemm(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda,
               int64_t ldb, int64_t ldc, int64_t stride_a, int64_t stride_b,
               int64_t flags, const void *A, const void *B, void *C,
               int64_t batch) {
  xo_fused_brgemm(dtype, m, n, k, lda, ldb, ldc, stride_a, stride_b, flags, 0, 0,
                  0, 0, A, B, C, NULL, batch);
}

/* gemm == brgemm with one batch (XsmmRunnerUtils.cpp:79-93) */
void xo_gemm(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda,
             int64_t ldb, int64_t ldc, int64_t flags, const void *A,
             const void *B, void *C) {
  xo_brgemm(dtype, m, n, k, lda, ldb, ldc, 0, 0, flags, A, B, C, 1);
}

/* [m,n] are the OUTPUT dims, except transpose / vnni_2 where they are the INPUT
 * dims (ConvertLinalgToXsmm.cpp:147-148, :1060-1074). */
int xo_unary(int64_t kind, int64_t dtype, int64_t m, int64_t n, int64_t ldi,
             int64_t ldo, int64_t flags, const void *in, void *out) {
  const size_t es = dtype == XO_F32 ? 4 : 2;
  if (kind == 29) { /* transpose: out[j][i] = in[i][j], bit copy */
    for (int64_t i = 0; i < m; ++i)
      for (int64_t j = 0; j < n; ++j)
        memcpy((char *)out + (size_t)(j * ldo + i) * es,
               (const char *)in + (size_t)(i * ldi + j) * es, es);
    return 0;
  }
  if (kind == 28) { /* norm -> VNNI2: in m x n (K x N) -> out [m/2][n][2] */
    if (dtype != XO_BF16 || (m % 2) != 0)
      return -1;
    const uint16_t *src = (const uint16_t *)in;
    uint16_t *dst = (uint16_t *)out;
    for (int64_t p = 0; p < m; ++p)
      for (int64_t j = 0; j < n; ++j)
        dst[((p / 2) * ldo + j) * 2 + (p % 2)] = src[p * ldi + j];
    return 0;
  }
  if (kind == 32 || kind == 1032) { /* norm -> VNNI4 ([m/4][n][4]) and its inverse (extension), m x n = the flat dims */
    if (dtype != XO_BF16 || (m % 4) != 0)
      return -1;
    const uint16_t *src = (const uint16_t *)in;
    uint16_t *dst = (uint16_t *)out;
    for (int64_t p = 0; p < m; ++p)
      for (int64_t j = 0; j < n; ++j) {
        if (kind == 32)
          dst[((p / 4) * ldo + j) * 4 + (p % 4)] = src[p * ldi + j];
        else
          dst[p * ldo + j] = src[((p / 4) * ldi + j) * 4 + (p % 4)];
      }
    return 0;
  }
  if (kind == 1028) { /* extension: VNNI2 -> norm, m x n are the OUTPUT dims */
    if (dtype != XO_BF16 || (m % 2) != 0)
      return -1;
    const uint16_t *src = (const uint16_t *)in;
    uint16_t *dst = (uint16_t *)out;
    for (int64_t p = 0; p < m; ++p)
      for (int64_t j = 0; j < n; ++j)
        dst[p * ldo + j] = src[((p / 2) * ldi + j) * 2 + (p % 2)];
    return 0;
  }
  if (kind != 1 && kind != 2 && kind != 5)
    return -1;
  for (int64_t i = 0; i < m; ++i)
    for (int64_t j = 0; j < n; ++j) {
      if (kind == 2) { /* zero: +0 bits, input ignored (libxsmm XOR) */
        memset((char *)out + (size_t)(i * ldo + j) * es, 0, es);
        continue;
      }
      int64_t sidx = flags == 0   ? i * ldi + j
                     : flags == 2 ? i * ldi /* bcast_row: in[i][0], ldi == 1 */
                     : flags == 4 ? j       /* bcast_col: in[0][j] */
                                  : 0;      /* bcast_scalar */
      if (kind == 1) { /* identity: bit copy */
        memcpy((char *)out + (size_t)(i * ldo + j) * es,
               (const char *)in + (size_t)sidx * es, es);
      } else { /* relu in f32, rounded back */
        st(dtype, out, i * ldo + j, relu(ld(dtype, in, sidx)));
      }
    }
  return 0;
}

int xo_binary(int64_t kind, int64_t dtype, int64_t m, int64_t n, int64_t ldl,
              int64_t ldr, int64_t ldo, int64_t flags, const void *lhs,
              const void *rhs, void *out) {
  if (kind < 1 || kind > 4)
    return -1;
  for (int64_t i = 0; i < m; ++i)
    for (int64_t j = 0; j < n; ++j) {
      int64_t li = (flags & 1) ? i * ldl : (flags & 4) ? j : (flags & 16) ? 0 : i * ldl + j;
      int64_t ri = (flags & 2) ? i * ldr : (flags & 8) ? j : (flags & 32) ? 0 : i * ldr + j;
      float l = ld(dtype, lhs, li), r = ld(dtype, rhs, ri);
      float v = kind == 1 ? l + r : kind == 2 ? l * r : kind == 3 ? l - r : l / r;
      st(dtype, out, i * ldo + j, v);
    }
  return 0;
}
