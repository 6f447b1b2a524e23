// eltwise.cu - unary / binary / transform TPPs as HBM-bandwidth kernels.
//
// Replaces the libxsmm meltw kernels the reference dispatches in
// runtime/Xsmm/XsmmRunnerUtils.cpp:142-211 (unary: identity, zero, relu,
// transpose, vnni_2; binary: add, mul, sub, div; all with the broadcast modes of
// lib/TPP/Dialect/Xsmm/XsmmUtils.cpp:90-252).
//
// These ops have no data reuse: the design rule is 16-byte vector accesses,
// fully coalesced on both sides, several independent loads in flight per thread
// (the first version had one and reached 41-62 % of the HBM copy rate), no integer
// division in the address path, grids of many small CTAs. bf16 arithmetic is done
// in f32 and rounded once (comp_type F32, XsmmRunnerUtils.cpp:161-164,193-195); pure
// data movement (identity, zero, transpose, vnni) moves bits.
#include "common.cuh"
#include "kernels.h"

namespace tpp {

namespace {

constexpr int kThreads = 256;
constexpr int kNumSMs = 148;

__device__ __forceinline__ float apply_op(int op, float a, float b) {
  switch (op) {
  case kOpRelu: return relu_f32(a);
  case kOpAdd: return a + b;
  case kOpMul: return a * b;
  case kOpSub: return a - b;
  case kOpDiv: return a / b;
  default: return a;
  }
}

// ---- scalar (any alignment) path -------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) eltwise_scalar_kernel(EltwiseArgs a) {
  const int64_t total = a.m * a.n;
  const T *in0 = static_cast<const T *>(a.in0);
  const T *in1 = static_cast<const T *>(a.in1);
  T *out = static_cast<T *>(a.out);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / a.n, j = idx - i * a.n;
    if (a.op == kOpZero) {
      out[i * a.ldo + j] = T(0);
      continue;
    }
    const int64_t i0 = a.mode0 == kBcastNone  ? i * a.ld0 + j
                       : a.mode0 == kBcastRow ? i * a.ld0
                       : a.mode0 == kBcastCol ? j
                                              : 0;
    if (a.mode0 == kBcastImm) { // scalar passed by value, rounded once to the element type
      const float r = apply_op(a.op == kOpIdentity ? kOpIdentity : a.op, a.imm, 0.f);
      if constexpr (sizeof(T) == 4) out[i * a.ldo + j] = r; else out[i * a.ldo + j] = f32_to_bf16_bits(r);
      continue;
    }
    if (a.op == kOpIdentity) { // bit copy
      out[i * a.ldo + j] = in0[i0];
      continue;
    }
    float x, y = 0.f;
    if constexpr (sizeof(T) == 4) x = in0[i0]; else x = bf16_bits_to_f32(in0[i0]);
    if (a.op >= kOpAdd) {
      const int64_t i1 = a.mode1 == kBcastNone  ? i * a.ld1 + j
                         : a.mode1 == kBcastRow ? i * a.ld1
                         : a.mode1 == kBcastCol ? j
                                                : 0;
      if constexpr (sizeof(T) == 4) y = in1[i1]; else y = bf16_bits_to_f32(in1[i1]);
    }
    const float r = apply_op(a.op, x, y);
    if constexpr (sizeof(T) == 4) out[i * a.ldo + j] = r; else out[i * a.ldo + j] = f32_to_bf16_bits(r);
  }
}

// ---- 16-byte vector path ------------------------------------------------------
// A CTA is 32 column-vectors (32 x 16 B = 512 contiguous bytes per row) x 8 rows; every thread handles
// R rows (stride 8), so R independent 16-byte loads per operand are in flight before the first use.

template <typename T, int VEC> __device__ __forceinline__ void unpack16(const uint4 &raw, float (&x)[VEC]) {
  if constexpr (sizeof(T) == 4) {
    x[0] = __uint_as_float(raw.x); x[1] = __uint_as_float(raw.y);
    x[2] = __uint_as_float(raw.z); x[3] = __uint_as_float(raw.w);
  } else {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      x[2 * q] = __uint_as_float(w[q] << 16);
      x[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
    }
  }
}

template <typename T> __device__ __forceinline__ uint4 splat16(T s) {
  uint32_t bits;
  if constexpr (sizeof(T) == 4) bits = __float_as_uint(s); else bits = (uint32_t)s | ((uint32_t)s << 16);
  return make_uint4(bits, bits, bits, bits);
}

// raw 16 bytes of operand `base` for row i, first column j
template <typename T>
__device__ __forceinline__ uint4 load16(const T *base, int mode, int64_t ld, int64_t i, int64_t j) {
  if (mode == kBcastNone) return *reinterpret_cast<const uint4 *>(base + i * ld + j);
  if (mode == kBcastCol) return *reinterpret_cast<const uint4 *>(base + j);
  return splat16<T>(mode == kBcastRow ? base[i * ld] : base[0]);
}

// 2-D grid of (32 column-vectors) x (8 * rows_per_thread rows) tiles; y is capped, the kernel strides over it
inline dim3 grid2d(int64_t nvec, int64_t rows, int rows_per_cta) {
  int64_t gx = (nvec + 31) / 32, gy = (rows + rows_per_cta - 1) / rows_per_cta;
  const int64_t cap = (int64_t)kNumSMs * 32;
  if (gx * gy > cap) gy = cap / gx > 0 ? cap / gx : 1;
  if (gy > 65535) gy = 65535;
  return dim3((unsigned)gx, (unsigned)gy, 1);
}

// NIN = number of tensor inputs (0: zero, 1: identity / relu, 2: binary), R = rows per thread. Specialising on
// NIN keeps the register count low enough for 5-6 resident CTAs per SM (about 100 KiB of loads in flight per SM).
// INV (binary only): 1 / 2 = operand 0 / 1 does not depend on the row (bcast_col: the bias vector, or a scalar);
// it is loaded once per thread instead of once per row, which leaves registers for R = 4 rows in flight.
template <typename T, int VEC, int NIN, int R, int INV>
__global__ void __launch_bounds__(kThreads, (NIN == 2 && INV == 0) ? 4 : 5) eltwise_vec_kernel(EltwiseArgs a) {
  const int64_t nv = a.n / VEC;
  const int64_t cv = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  if (cv >= nv) return;
  const int64_t j = cv * VEC;
  const int ty = threadIdx.x >> 5;
  const T *in0 = static_cast<const T *>(a.in0);
  const T *in1 = static_cast<const T *>(a.in1);
  T *out = static_cast<T *>(a.out);
  uint4 inv = make_uint4(0, 0, 0, 0);
  if constexpr (INV == 1) inv = load16<T>(in0, a.mode0, a.ld0, 0, j);
  if constexpr (INV == 2) inv = load16<T>(in1, a.mode1, a.ld1, 0, j);
  for (int64_t r0 = (int64_t)blockIdx.y * (8 * R) + ty; r0 < a.m; r0 += (int64_t)gridDim.y * (8 * R)) {
    uint4 raw0[R], raw1[INV == 0 ? R : 1];
    if constexpr (NIN >= 1) {
#pragma unroll
      for (int u = 0; u < R; ++u) {
        const int64_t i = r0 + 8 * u;
        if (i < a.m) {
          // with an invariant operand, raw0[] carries the streaming one
          if constexpr (INV == 1) raw0[u] = load16<T>(in1, a.mode1, a.ld1, i, j);
          else raw0[u] = load16<T>(in0, a.mode0, a.ld0, i, j);
          if constexpr (NIN == 2 && INV == 0) raw1[u] = load16<T>(in1, a.mode1, a.ld1, i, j);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int64_t i = r0 + 8 * u;
      if (i >= a.m) continue;   // (not break: keeps the loop fully unrolled and raw*[] in registers)
      uint4 o;
      if constexpr (NIN == 0) {
        o = make_uint4(0, 0, 0, 0);
      } else {
        if (NIN == 1 && a.op == kOpIdentity) {
          o = raw0[u];
        } else {
          float x[VEC], y[VEC], r[VEC];
          if constexpr (INV == 1) {          // operand 0 is the invariant one, raw0[] holds operand 1
            unpack16<T, VEC>(inv, x);
            unpack16<T, VEC>(raw0[u], y);
          } else {
            unpack16<T, VEC>(raw0[u], x);
            if constexpr (NIN == 2) unpack16<T, VEC>(INV == 2 ? inv : raw1[u], y);
          }
#pragma unroll
          for (int q = 0; q < VEC; ++q) r[q] = NIN == 2 ? apply_op(a.op, x[q], y[q]) : relu_f32(x[q]);
          if constexpr (sizeof(T) == 4)
            o = make_uint4(__float_as_uint(r[0]), __float_as_uint(r[1]), __float_as_uint(r[2]), __float_as_uint(r[3]));
          else
            o = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]),
                           pack_bf16x2(r[6], r[7]));
        }
      }
      *reinterpret_cast<uint4 *>(out + i * a.ldo + j) = o;
    }
  }
}

template <typename T, int VEC> void launch_vec(const EltwiseArgs &a, dim3 (*g2)(int64_t, int64_t, int), cudaStream_t s) {
  const bool inv0 = a.mode0 == kBcastCol || a.mode0 == kBcastScalar;
  const bool inv1 = a.mode1 == kBcastCol || a.mode1 == kBcastScalar;
  if (a.op == kOpZero)
    eltwise_vec_kernel<T, VEC, 0, 4, 0><<<g2(a.n / VEC, a.m, 32), kThreads, 0, s>>>(a);
  else if (a.op < kOpAdd)
    eltwise_vec_kernel<T, VEC, 1, 4, 0><<<g2(a.n / VEC, a.m, 32), kThreads, 0, s>>>(a);
  else if (inv0)
    eltwise_vec_kernel<T, VEC, 2, 4, 1><<<g2(a.n / VEC, a.m, 32), kThreads, 0, s>>>(a);
  else if (inv1)
    eltwise_vec_kernel<T, VEC, 2, 4, 2><<<g2(a.n / VEC, a.m, 32), kThreads, 0, s>>>(a);
  else
    eltwise_vec_kernel<T, VEC, 2, 2, 0><<<g2(a.n / VEC, a.m, 16), kThreads, 0, s>>>(a);
}

inline int grid_for(int64_t work_items) {
  int64_t blocks = (work_items + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)kNumSMs * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

inline bool operand_vec_ok(const void *p, int mode, int64_t ld, int vec) {
  if (mode == kBcastNone) return aligned16(p) && (ld % vec) == 0;
  if (mode == kBcastCol) return aligned16(p);
  return true;
}

// ---- transpose --------------------------------------------------------------
// Generic: 64x64 element tile through shared memory, element-wise (any size / alignment).
template <typename T>
__global__ void __launch_bounds__(256) transpose_kernel(const T *__restrict__ in, T *__restrict__ out, int64_t m,
                                                        int64_t n, int64_t ldi, int64_t ldo) {
  constexpr int TILE = 64;
  __shared__ T tile[TILE][TILE + (sizeof(T) == 2 ? 2 : 1)];
  const int64_t tiles_n = (n + TILE - 1) / TILE;
  const int64_t tiles_m = (m + TILE - 1) / TILE;
  for (int64_t t = blockIdx.x; t < tiles_m * tiles_n; t += gridDim.x) {
    const int64_t i0 = (t / tiles_n) * TILE, j0 = (t % tiles_n) * TILE;
    const int tx = threadIdx.x % TILE, ty = threadIdx.x / TILE; // 64 x 4
#pragma unroll 4
    for (int r = ty; r < TILE; r += 4) {
      const int64_t i = i0 + r, j = j0 + tx;
      if (i < m && j < n) tile[r][tx] = in[i * ldi + j];
    }
    __syncthreads();
#pragma unroll 4
    for (int r = ty; r < TILE; r += 4) {
      const int64_t j = j0 + r, i = i0 + tx; // out row j, col i
      if (i < m && j < n) out[j * ldo + i] = tile[tx][r];
    }
    __syncthreads();
  }
}

// bf16, even m / n / ld, 4-byte aligned: every global and shared access is 32 bits wide. A thread loads the
// 2x2 block (rows 2rp,2rp+1 x cols 2cp,2cp+1) as two words, transposes it in registers with two byte-permutes
// and stores the two words of the transposed tile; the write phase reads full 128-byte output rows.
__global__ void __launch_bounds__(256) transpose_bf16_pair_kernel(const uint16_t *__restrict__ in,
                                                                  uint16_t *__restrict__ out, int64_t m, int64_t n,
                                                                  int64_t ldi, int64_t ldo) {
  __shared__ uint32_t tile[64][33];   // [output row within tile (input column)][input row pair]
  const int64_t tiles_n = (n + 63) / 64;
  const int64_t i0 = (int64_t)(blockIdx.x / tiles_n) * 64, j0 = (int64_t)(blockIdx.x % tiles_n) * 64;
  const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
  uint32_t w0[4], w1[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {   // all loads first
    const int rp = y + 8 * u;
    const int64_t i = i0 + 2 * rp, j = j0 + 2 * x;
    w0[u] = w1[u] = 0;
    if (i < m && j < n) {
      w0[u] = *reinterpret_cast<const uint32_t *>(in + i * ldi + j);
      w1[u] = *reinterpret_cast<const uint32_t *>(in + (i + 1) * ldi + j);
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int rp = y + 8 * u;
    tile[2 * x][rp] = __byte_perm(w0[u], w1[u], 0x5410);       // (in[2rp][2x],   in[2rp+1][2x])
    tile[2 * x + 1][rp] = __byte_perm(w0[u], w1[u], 0x7632);   // (in[2rp][2x+1], in[2rp+1][2x+1])
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int c = y + 8 * u;                 // output row within the tile
    const int64_t jo = j0 + c, io = i0 + 2 * x;
    if (jo < n && io < m) *reinterpret_cast<uint32_t *>(out + jo * ldo + io) = tile[c][x];
  }
}

// ---- VNNI-2 pack / unpack -----------------------------------------------------
// pack:  out[((p/2)*ldo + j)*2 + p%2] = in[p*ldi + j]   (p<m=K, j<n=N)
// A thread owns 8 columns of PAIRS row pairs (stride 8 pairs): 2*PAIRS 16-byte loads in flight, then
// one 32-byte contiguous interleaved store per pair.
constexpr int PAIRS = 2;

__global__ void __launch_bounds__(kThreads) vnni2_pack_vec_kernel(const uint16_t *__restrict__ in,
                                                                 uint16_t *__restrict__ out, int64_t m, int64_t n,
                                                                 int64_t ldi, int64_t ldo) {
  const int64_t nv = n / 8, np = m / 2;
  const int64_t cv = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  if (cv >= nv) return;
  const int64_t j = cv * 8;
  const int ty = threadIdx.x >> 5;
  for (int64_t q0 = (int64_t)blockIdx.y * (8 * PAIRS) + ty; q0 < np; q0 += (int64_t)gridDim.y * (8 * PAIRS)) {
    uint4 r0[PAIRS], r1[PAIRS];
#pragma unroll
    for (int u = 0; u < PAIRS; ++u) {
      const int64_t q = q0 + 8 * u;
      if (q < np) {
        r0[u] = *reinterpret_cast<const uint4 *>(in + (2 * q) * ldi + j);
        r1[u] = *reinterpret_cast<const uint4 *>(in + (2 * q + 1) * ldi + j);
      }
    }
#pragma unroll
    for (int u = 0; u < PAIRS; ++u) {
      const int64_t q = q0 + 8 * u;
      if (q >= np) continue;
      uint4 o0, o1;
      o0.x = __byte_perm(r0[u].x, r1[u].x, 0x5410); o0.y = __byte_perm(r0[u].x, r1[u].x, 0x7632);
      o0.z = __byte_perm(r0[u].y, r1[u].y, 0x5410); o0.w = __byte_perm(r0[u].y, r1[u].y, 0x7632);
      o1.x = __byte_perm(r0[u].z, r1[u].z, 0x5410); o1.y = __byte_perm(r0[u].z, r1[u].z, 0x7632);
      o1.z = __byte_perm(r0[u].w, r1[u].w, 0x5410); o1.w = __byte_perm(r0[u].w, r1[u].w, 0x7632);
      uint4 *dst = reinterpret_cast<uint4 *>(out + (q * ldo + j) * 2);
      dst[0] = o0;
      dst[1] = o1;
    }
  }
}

__global__ void __launch_bounds__(kThreads) vnni2_pack_scalar_kernel(const uint16_t *__restrict__ in,
                                                                    uint16_t *__restrict__ out, int64_t m, int64_t n,
                                                                    int64_t ldi, int64_t ldo) {
  const int64_t total = m * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx / n, j = idx - p * n;
    out[((p / 2) * ldo + j) * 2 + (p % 2)] = in[p * ldi + j];
  }
}

// unpack: out[p*ldo + j] = in[((p/2)*ldi + j)*2 + p%2]
__global__ void __launch_bounds__(kThreads) vnni2_unpack_vec_kernel(const uint16_t *__restrict__ in,
                                                                   uint16_t *__restrict__ out, int64_t m, int64_t n,
                                                                   int64_t ldi, int64_t ldo) {
  const int64_t nv = n / 8, np = m / 2;
  const int64_t cv = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  if (cv >= nv) return;
  const int64_t j = cv * 8;
  const int ty = threadIdx.x >> 5;
  for (int64_t q0 = (int64_t)blockIdx.y * (8 * PAIRS) + ty; q0 < np; q0 += (int64_t)gridDim.y * (8 * PAIRS)) {
    uint4 a[PAIRS], b[PAIRS];
#pragma unroll
    for (int u = 0; u < PAIRS; ++u) {
      const int64_t q = q0 + 8 * u;
      if (q < np) {
        const uint4 *src = reinterpret_cast<const uint4 *>(in + (q * ldi + j) * 2);
        a[u] = src[0];
        b[u] = src[1];
      }
    }
#pragma unroll
    for (int u = 0; u < PAIRS; ++u) {
      const int64_t q = q0 + 8 * u;
      if (q >= np) continue;
      uint4 r0, r1;
      r0.x = __byte_perm(a[u].x, a[u].y, 0x5410); r1.x = __byte_perm(a[u].x, a[u].y, 0x7632);
      r0.y = __byte_perm(a[u].z, a[u].w, 0x5410); r1.y = __byte_perm(a[u].z, a[u].w, 0x7632);
      r0.z = __byte_perm(b[u].x, b[u].y, 0x5410); r1.z = __byte_perm(b[u].x, b[u].y, 0x7632);
      r0.w = __byte_perm(b[u].z, b[u].w, 0x5410); r1.w = __byte_perm(b[u].z, b[u].w, 0x7632);
      *reinterpret_cast<uint4 *>(out + (2 * q) * ldo + j) = r0;
      *reinterpret_cast<uint4 *>(out + (2 * q + 1) * ldo + j) = r1;
    }
  }
}

__global__ void __launch_bounds__(kThreads) vnni2_unpack_scalar_kernel(const uint16_t *__restrict__ in,
                                                                      uint16_t *__restrict__ out, int64_t m,
                                                                      int64_t n, int64_t ldi, int64_t ldo) {
  const int64_t total = m * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx / n, j = idx - p * n;
    out[p * ldo + j] = in[((p / 2) * ldi + j) * 2 + (p % 2)];
  }
}

// ---- VNNI-4 pack / unpack ([K][N] <-> [K/4][N][4]) -------------------------------------------------------------
// pack: out[((p/4)*ldo + j)*4 + p%4] = in[p*ldi + j]. A thread owns 8 columns of one row quad: four 16-byte loads (one
// per row), four 16-byte stores of 64 contiguous bytes (8 columns x 4 k).
__global__ void __launch_bounds__(kThreads) vnni4_pack_vec_kernel(const uint16_t *__restrict__ in,
                                                                 uint16_t *__restrict__ out, int64_t m, int64_t n,
                                                                 int64_t ldi, int64_t ldo) {
  const int64_t nv = n / 8, nq = m / 4;
  const int64_t cv = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  if (cv >= nv) return;
  const int64_t j = cv * 8;
  const int ty = threadIdx.x >> 5;
  for (int64_t q = (int64_t)blockIdx.y * 8 + ty; q < nq; q += (int64_t)gridDim.y * 8) {
    uint4 r[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) r[v] = *reinterpret_cast<const uint4 *>(in + (4 * q + v) * ldi + j);
    const uint32_t w[4][4] = {{r[0].x, r[0].y, r[0].z, r[0].w}, {r[1].x, r[1].y, r[1].z, r[1].w},
                              {r[2].x, r[2].y, r[2].z, r[2].w}, {r[3].x, r[3].y, r[3].z, r[3].w}};
    uint4 *dst = reinterpret_cast<uint4 *>(out + (q * ldo + j) * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {   // word i of a row holds columns 2i (low half) and 2i + 1 (high half)
      uint4 o;
      o.x = __byte_perm(w[0][i], w[1][i], 0x5410); o.y = __byte_perm(w[2][i], w[3][i], 0x5410);   // column 2i: k 0,1 | k 2,3
      o.z = __byte_perm(w[0][i], w[1][i], 0x7632); o.w = __byte_perm(w[2][i], w[3][i], 0x7632);   // column 2i + 1
      dst[i] = o;
    }
  }
}

__global__ void __launch_bounds__(kThreads) vnni4_unpack_vec_kernel(const uint16_t *__restrict__ in,
                                                                   uint16_t *__restrict__ out, int64_t m, int64_t n,
                                                                   int64_t ldi, int64_t ldo) {
  const int64_t nv = n / 8, nq = m / 4;
  const int64_t cv = (int64_t)blockIdx.x * 32 + (threadIdx.x & 31);
  if (cv >= nv) return;
  const int64_t j = cv * 8;
  const int ty = threadIdx.x >> 5;
  for (int64_t q = (int64_t)blockIdx.y * 8 + ty; q < nq; q += (int64_t)gridDim.y * 8) {
    const uint4 *src = reinterpret_cast<const uint4 *>(in + (q * ldi + j) * 4);
    uint4 a[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = src[i];   // a[i] = columns 2i, 2i + 1: (k0 k1 | k2 k3) each
    uint32_t row[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      row[0][i] = __byte_perm(a[i].x, a[i].z, 0x5410);   // k0 of column 2i | k0 of column 2i + 1
      row[1][i] = __byte_perm(a[i].x, a[i].z, 0x7632);
      row[2][i] = __byte_perm(a[i].y, a[i].w, 0x5410);
      row[3][i] = __byte_perm(a[i].y, a[i].w, 0x7632);
    }
#pragma unroll
    for (int v = 0; v < 4; ++v)
      *reinterpret_cast<uint4 *>(out + (4 * q + v) * ldo + j) = make_uint4(row[v][0], row[v][1], row[v][2], row[v][3]);
  }
}

__global__ void __launch_bounds__(kThreads) vnni4_scalar_kernel(const uint16_t *__restrict__ in, uint16_t *__restrict__ out,
                                                               int64_t m, int64_t n, int64_t ldi, int64_t ldo, int unpack) {
  const int64_t total = m * n;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = idx / n, j = idx - p * n;
    if (unpack) out[p * ldo + j] = in[((p / 4) * ldi + j) * 4 + (p % 4)];
    else out[((p / 4) * ldo + j) * 4 + (p % 4)] = in[p * ldi + j];
  }
}

} // namespace

void launch_eltwise(const EltwiseArgs &a_in, cudaStream_t stream) {
  if (a_in.m <= 0 || a_in.n <= 0) return;
  EltwiseArgs a = a_in;
  const bool f32 = a.dtype == kF32;
  const int vec = f32 ? 4 : 8;
  // fully contiguous operands (no row/col broadcast): re-shape to rows of 8192 elements so that tall-skinny
  // tensors still give every lane a 16-byte vector
  auto contiguous = [&](int mode, int64_t ld) { return (mode == kBcastNone && ld == a.n) || mode >= kBcastScalar; };
  if (a.ldo == a.n && (a.op == kOpZero || contiguous(a.mode0, a.ld0)) && (a.op < kOpAdd || contiguous(a.mode1, a.ld1))) {
    const int64_t total = a.m * a.n;
    if (a.n < 2048 && total % 8192 == 0) {
      a.n = 8192; a.m = total / 8192;
      a.ldo = 8192;
      if (a.mode0 == kBcastNone) a.ld0 = 8192;
      if (a.mode1 == kBcastNone) a.ld1 = 8192;
    }
  }
  bool vec_ok = (a.n % vec) == 0 && aligned16(a.out) && (a.ldo % vec) == 0;
  if (a.op != kOpZero) vec_ok = vec_ok && a.mode0 != kBcastImm && operand_vec_ok(a.in0, a.mode0, a.ld0, vec);
  if (a.op >= kOpAdd) vec_ok = vec_ok && operand_vec_ok(a.in1, a.mode1, a.ld1, vec);
  if (vec_ok) {
    if (f32) launch_vec<float, 4>(a, grid2d, stream);
    else launch_vec<uint16_t, 8>(a, grid2d, stream);
  } else {
    const int grid = grid_for(a.m * a.n);
    if (f32) eltwise_scalar_kernel<float><<<grid, kThreads, 0, stream>>>(a);
    else eltwise_scalar_kernel<uint16_t><<<grid, kThreads, 0, stream>>>(a);
  }
  TPP_CUDA_CHECK(cudaGetLastError());
}

void launch_transpose(const void *in, void *out, int64_t m, int64_t n, int64_t ldi, int64_t ldo, int es,
                      cudaStream_t stream) {
  if (m <= 0 || n <= 0) return;
  const int64_t tiles = ((m + 63) / 64) * ((n + 63) / 64);
  const bool pair_ok = es == 2 && (m % 2) == 0 && (n % 2) == 0 && (ldi % 2) == 0 && (ldo % 2) == 0 &&
                       (reinterpret_cast<uintptr_t>(in) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 3) == 0 &&
                       tiles < (1ll << 31);
  if (pair_ok) {
    transpose_bf16_pair_kernel<<<(unsigned)tiles, 256, 0, stream>>>(static_cast<const uint16_t *>(in),
                                                                    static_cast<uint16_t *>(out), m, n, ldi, ldo);
  } else {
    const int grid = (int)(tiles < (int64_t)kNumSMs * 8 ? tiles : (int64_t)kNumSMs * 8);
    if (es == 4)
      transpose_kernel<uint32_t><<<grid, 256, 0, stream>>>(static_cast<const uint32_t *>(in),
                                                           static_cast<uint32_t *>(out), m, n, ldi, ldo);
    else
      transpose_kernel<uint16_t><<<grid, 256, 0, stream>>>(static_cast<const uint16_t *>(in),
                                                           static_cast<uint16_t *>(out), m, n, ldi, ldo);
  }
  TPP_CUDA_CHECK(cudaGetLastError());
}

// ---- batched tile moves (block-layout pack / unpack as ONE kernel) -----------------------------------
// tensor.pack / tensor.unpack reach the ABI as one xsmm_unary_invoke (identity, or transpose) per tile with the same
// descriptor (lib/TPP/Transforms/LowerPacksAndUnpacks.cpp:143-250 tiles the pack by one outer tile,
// ConvertLinalgToXsmm turns each tile's copy / transpose into a unary TPP): hundreds of 2-4 KiB moves. During graph
// capture the runtime collects such runs and launches this kernel once: CTA t moves tile t, (in, out) pointers come from a
// device table.
template <typename V>
__global__ void __launch_bounds__(256) tile_copy_batch_kernel(const TilePtrs *__restrict__ tiles, int64_t m, int64_t nv,
                                                              int64_t ldi_v, int64_t ldo_v) {
  const TilePtrs t = tiles[blockIdx.x];
  const V *__restrict__ in = static_cast<const V *>(t.in);
  V *__restrict__ out = static_cast<V *>(t.out);
  const int64_t total = m * nv;
  for (int64_t e = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.y * blockDim.x) {
    const int64_t r = e / nv, c = e - r * nv;
    out[r * ldo_v + c] = in[r * ldi_v + c];
  }
}

// 16-byte form for small tiles (a 32 x 32 bf16 tile is 128 vectors): one CTA per tile leaves half of its threads idle and
// pays a CTA launch per 2 KiB (45 % of the copy bandwidth at 4096^2, round 1). Here the (tile, row, vector) space is
// flattened: a CTA takes 1024 consecutive vectors (several tiles), every thread issues its four independent loads
// before the first store.
__global__ void __launch_bounds__(256) tile_copy_flat_kernel(const TilePtrs *__restrict__ tiles, uint32_t num_tiles, uint32_t m,
                                                             uint32_t nv, int64_t ldi_v, int64_t ldo_v) {
  constexpr int U = 4;
  const uint32_t per_tile = m * nv;
  const uint64_t total = (uint64_t)num_tiles * per_tile;
  const uint64_t base = (uint64_t)blockIdx.x * (256 * U) + threadIdx.x;
  uint4 v[U];
  uint4 *dst[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const uint64_t g = base + (uint64_t)u * 256;
    dst[u] = nullptr;
    if (g < total) {
      const uint32_t t = (uint32_t)(g / per_tile), e = (uint32_t)(g - (uint64_t)t * per_tile);
      const uint32_t r = e / nv, c = e - r * nv;
      const TilePtrs tp = tiles[t];
      v[u] = __ldg(static_cast<const uint4 *>(tp.in) + (int64_t)r * ldi_v + c);
      dst[u] = static_cast<uint4 *>(tp.out) + (int64_t)r * ldo_v + c;
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (dst[u]) *dst[u] = v[u];
}

template <typename T>
__global__ void __launch_bounds__(256) tile_transpose_batch_kernel(const TilePtrs *__restrict__ tiles, int64_t m, int64_t n,
                                                                   int64_t ldi, int64_t ldo) {
  constexpr int TILE = 32;
  __shared__ T sm[TILE][TILE + 1];
  const TilePtrs t = tiles[blockIdx.x];
  const T *__restrict__ in = static_cast<const T *>(t.in);
  T *__restrict__ out = static_cast<T *>(t.out);
  const int64_t tiles_n = (n + TILE - 1) / TILE, tiles_m = (m + TILE - 1) / TILE;
  const int tx = threadIdx.x % TILE, ty = threadIdx.x / TILE;   // 32 x 8
  for (int64_t b = blockIdx.y; b < tiles_m * tiles_n; b += gridDim.y) {
    const int64_t i0 = (b / tiles_n) * TILE, j0 = (b % tiles_n) * TILE;
#pragma unroll
    for (int r = ty; r < TILE; r += 8) {
      const int64_t i = i0 + r, j = j0 + tx;
      if (i < m && j < n) sm[r][tx] = in[i * ldi + j];
    }
    __syncthreads();
#pragma unroll
    for (int r = ty; r < TILE; r += 8) {
      const int64_t j = j0 + r, i = i0 + tx;   // out row j, col i
      if (i < m && j < n) out[j * ldo + i] = sm[tx][r];
    }
    __syncthreads();
  }
}

void launch_tile_batch(const TilePtrs *dev_tiles, int64_t num_tiles, bool transpose, int64_t m, int64_t n, int64_t ldi,
                       int64_t ldo, int es, bool vec16_ok, cudaStream_t stream) {
  if (num_tiles <= 0 || m <= 0 || n <= 0) return;
  if (transpose) {
    const int64_t sub = ((m + 31) / 32) * ((n + 31) / 32);
    dim3 grid((unsigned)num_tiles, (unsigned)(sub < 64 ? sub : 64));
    if (es == 4) tile_transpose_batch_kernel<uint32_t><<<grid, 256, 0, stream>>>(dev_tiles, m, n, ldi, ldo);
    else tile_transpose_batch_kernel<uint16_t><<<grid, 256, 0, stream>>>(dev_tiles, m, n, ldi, ldo);
  } else if (vec16_ok) {
    const int64_t per = 16 / es, nv = n / per, chunks = (m * nv + 255) / 256;
    if (m * nv <= 4096 && num_tiles * m * nv < (1ll << 40) && num_tiles < (1ll << 31)) {
      const int64_t total = num_tiles * m * nv;
      tile_copy_flat_kernel<<<(unsigned)((total + 1023) / 1024), 256, 0, stream>>>(dev_tiles, (uint32_t)num_tiles, (uint32_t)m,
                                                                                   (uint32_t)nv, ldi / per, ldo / per);
    } else {
      dim3 grid((unsigned)num_tiles, (unsigned)(chunks < 64 ? chunks : 64));
      tile_copy_batch_kernel<uint4><<<grid, 256, 0, stream>>>(dev_tiles, m, nv, ldi / per, ldo / per);
    }
  } else {
    const int64_t chunks = (m * n + 255) / 256;
    dim3 grid((unsigned)num_tiles, (unsigned)(chunks < 64 ? chunks : 64));
    if (es == 4) tile_copy_batch_kernel<uint32_t><<<grid, 256, 0, stream>>>(dev_tiles, m, n, ldi, ldo);
    else tile_copy_batch_kernel<uint16_t><<<grid, 256, 0, stream>>>(dev_tiles, m, n, ldi, ldo);
  }
  TPP_CUDA_CHECK(cudaGetLastError());
}

void launch_vnni2_pack(const void *in, void *out, int64_t m, int64_t n, int64_t ldi, int64_t ldo,
                       cudaStream_t stream) {
  if (m <= 0 || n <= 0) return;
  const uint16_t *src = static_cast<const uint16_t *>(in);
  uint16_t *dst = static_cast<uint16_t *>(out);
  const bool vec_ok = (n % 8) == 0 && (ldi % 8) == 0 && (ldo % 4) == 0 && aligned16(in) && aligned16(out);
  if (vec_ok)
    vnni2_pack_vec_kernel<<<grid2d(n / 8, m / 2, 8 * PAIRS), kThreads, 0, stream>>>(src, dst, m, n, ldi, ldo);
  else
    vnni2_pack_scalar_kernel<<<grid_for(m * n), kThreads, 0, stream>>>(src, dst, m, n, ldi, ldo);
  TPP_CUDA_CHECK(cudaGetLastError());
}

void launch_vnni4_pack(const void *in, void *out, int64_t m, int64_t n, int64_t ldi, int64_t ldo, cudaStream_t stream) {
  if (m <= 0 || n <= 0) return;
  const uint16_t *src = static_cast<const uint16_t *>(in);
  uint16_t *dst = static_cast<uint16_t *>(out);
  if ((n % 8) == 0 && (ldi % 8) == 0 && (ldo % 2) == 0 && aligned16(in) && aligned16(out))
    vnni4_pack_vec_kernel<<<grid2d(n / 8, m / 4, 8), kThreads, 0, stream>>>(src, dst, m, n, ldi, ldo);
  else
    vnni4_scalar_kernel<<<grid_for(m * n), kThreads, 0, stream>>>(src, dst, m, n, ldi, ldo, 0);
}

void launch_vnni4_unpack(const void *in, void *out, int64_t m, int64_t n, int64_t ldi, int64_t ldo, cudaStream_t stream) {
  if (m <= 0 || n <= 0) return;
  const uint16_t *src = static_cast<const uint16_t *>(in);
  uint16_t *dst = static_cast<uint16_t *>(out);
  if ((n % 8) == 0 && (ldi % 2) == 0 && (ldo % 8) == 0 && aligned16(in) && aligned16(out))
    vnni4_unpack_vec_kernel<<<grid2d(n / 8, m / 4, 8), kThreads, 0, stream>>>(src, dst, m, n, ldi, ldo);
  else
    vnni4_scalar_kernel<<<grid_for(m * n), kThreads, 0, stream>>>(src, dst, m, n, ldi, ldo, 1);
}

void launch_vnni2_unpack(const void *in, void *out, int64_t m, int64_t n, int64_t ldi, int64_t ldo,
                         cudaStream_t stream) {
  if (m <= 0 || n <= 0) return;
  const uint16_t *src = static_cast<const uint16_t *>(in);
  uint16_t *dst = static_cast<uint16_t *>(out);
  const bool vec_ok = (n % 8) == 0 && (ldo % 8) == 0 && (ldi % 4) == 0 && aligned16(in) && aligned16(out);
  if (vec_ok)
    vnni2_unpack_vec_kernel<<<grid2d(n / 8, m / 2, 8 * PAIRS), kThreads, 0, stream>>>(src, dst, m, n, ldi, ldo);
  else
    vnni2_unpack_scalar_kernel<<<grid_for(m * n), kThreads, 0, stream>>>(src, dst, m, n, ldi, ldo);
  TPP_CUDA_CHECK(cudaGetLastError());
}

} // namespace tpp
