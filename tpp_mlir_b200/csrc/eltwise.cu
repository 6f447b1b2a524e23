// eltwise.cu - unary / binary / transform TPPs as HBM-bandwidth kernels.
//
// Replaces the libxsmm meltw kernels the reference dispatches in
// runtime/Xsmm/XsmmRunnerUtils.cpp:142-211 (unary: identity, zero, relu,
// transpose, vnni_2; binary: add, mul, sub, div; all with the broadcast modes of
// lib/TPP/Dialect/Xsmm/XsmmUtils.cpp:90-252).
//
// These ops have no data reuse: the design rule is 16-byte vector accesses,
// fully coalesced on both sides, enough bytes in flight per SM, grid sized in
// multiples of the SM count. bf16 arithmetic is done in f32 and rounded once
// (comp_type F32, XsmmRunnerUtils.cpp:161-164,193-195); pure data movement
// (identity, zero, transpose, vnni) moves bits.
#include "common.cuh"
#include "kernels.h"

namespace tpp {

namespace {

constexpr int kThreads = 256;
constexpr int kNumSMs = 148;

template <typename T> struct Vec16 { uint4 v; };

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

// ---- 16-byte vector path ----------------------------------------------------
// VEC elements per thread per access (8 bf16 or 4 f32).
template <typename T, int VEC>
__device__ __forceinline__ void load_operand(const T *base, int mode, int64_t ld, int64_t i, int64_t j,
                                             float (&x)[VEC], uint4 &raw) {
  if (mode == kBcastNone || mode == kBcastCol) {
    const T *p = mode == kBcastNone ? base + i * ld + j : base + j;
    raw = *reinterpret_cast<const uint4 *>(p);
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
  } else {
    const T s = mode == kBcastRow ? base[i * ld] : base[0];
    float f;
    uint32_t bits;
    if constexpr (sizeof(T) == 4) { f = s; bits = __float_as_uint(f); }
    else { f = bf16_bits_to_f32(s); bits = (uint32_t)s | ((uint32_t)s << 16); }
#pragma unroll
    for (int q = 0; q < VEC; ++q) x[q] = f;
    raw = make_uint4(bits, bits, bits, bits);
  }
}

template <typename T, int VEC>
__global__ void __launch_bounds__(kThreads) eltwise_vec_kernel(EltwiseArgs a) {
  const int64_t nv = a.n / VEC;
  const int64_t total = a.m * nv;
  const T *in0 = static_cast<const T *>(a.in0);
  const T *in1 = static_cast<const T *>(a.in1);
  T *out = static_cast<T *>(a.out);
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = idx / nv, j = (idx - i * nv) * VEC;
    uint4 *dst = reinterpret_cast<uint4 *>(out + i * a.ldo + j);
    if (a.op == kOpZero) {
      *dst = make_uint4(0, 0, 0, 0);
      continue;
    }
    float x[VEC], y[VEC];
    uint4 raw0, raw1;
    load_operand<T, VEC>(in0, a.mode0, a.ld0, i, j, x, raw0);
    if (a.op == kOpIdentity) {
      *dst = raw0;
      continue;
    }
    if (a.op >= kOpAdd) load_operand<T, VEC>(in1, a.mode1, a.ld1, i, j, y, raw1);
    float r[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) r[q] = apply_op(a.op, x[q], a.op >= kOpAdd ? y[q] : 0.f);
    uint4 o;
    if constexpr (sizeof(T) == 4) {
      o = make_uint4(__float_as_uint(r[0]), __float_as_uint(r[1]), __float_as_uint(r[2]), __float_as_uint(r[3]));
    } else {
      o = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]),
                     pack_bf16x2(r[6], r[7]));
    }
    *dst = o;
  }
}

inline int grid_for(int64_t work_items) {
  int64_t blocks = (work_items + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)kNumSMs * 16; // 8 resident 256-thread CTAs per SM x 2 waves, grid-stride beyond
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
// 64x64 element tile through shared memory; reads and writes are both coalesced
// along the contiguous dimension. T is the element type (uint16_t / uint32_t).
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

// ---- VNNI-2 pack / unpack -----------------------------------------------------
// pack:  out[((p/2)*ldo + j)*2 + p%2] = in[p*ldi + j]   (p<m=K, j<n=N)
// Each thread owns one row pair (2q, 2q+1) x 8 columns: two 16-byte loads, one
// 32-byte contiguous interleaved store.
__global__ void __launch_bounds__(kThreads) vnni2_pack_vec_kernel(const uint16_t *__restrict__ in,
                                                                 uint16_t *__restrict__ out, int64_t m, int64_t n,
                                                                 int64_t ldi, int64_t ldo) {
  const int64_t nv = n / 8, total = (m / 2) * nv;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = idx / nv, j = (idx - q * nv) * 8;
    const uint4 r0 = *reinterpret_cast<const uint4 *>(in + (2 * q) * ldi + j);
    const uint4 r1 = *reinterpret_cast<const uint4 *>(in + (2 * q + 1) * ldi + j);
    uint4 o0, o1;
    o0.x = __byte_perm(r0.x, r1.x, 0x5410); o0.y = __byte_perm(r0.x, r1.x, 0x7632);
    o0.z = __byte_perm(r0.y, r1.y, 0x5410); o0.w = __byte_perm(r0.y, r1.y, 0x7632);
    o1.x = __byte_perm(r0.z, r1.z, 0x5410); o1.y = __byte_perm(r0.z, r1.z, 0x7632);
    o1.z = __byte_perm(r0.w, r1.w, 0x5410); o1.w = __byte_perm(r0.w, r1.w, 0x7632);
    uint4 *dst = reinterpret_cast<uint4 *>(out + (q * ldo + j) * 2);
    dst[0] = o0;
    dst[1] = o1;
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
  const int64_t nv = n / 8, total = (m / 2) * nv;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = idx / nv, j = (idx - q * nv) * 8;
    const uint4 *src = reinterpret_cast<const uint4 *>(in + (q * ldi + j) * 2);
    const uint4 a = src[0], b = src[1];
    uint4 r0, r1;
    r0.x = __byte_perm(a.x, a.y, 0x5410); r1.x = __byte_perm(a.x, a.y, 0x7632);
    r0.y = __byte_perm(a.z, a.w, 0x5410); r1.y = __byte_perm(a.z, a.w, 0x7632);
    r0.z = __byte_perm(b.x, b.y, 0x5410); r1.z = __byte_perm(b.x, b.y, 0x7632);
    r0.w = __byte_perm(b.z, b.w, 0x5410); r1.w = __byte_perm(b.z, b.w, 0x7632);
    *reinterpret_cast<uint4 *>(out + (2 * q) * ldo + j) = r0;
    *reinterpret_cast<uint4 *>(out + (2 * q + 1) * ldo + j) = r1;
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

} // namespace

void launch_eltwise(const EltwiseArgs &a, cudaStream_t stream) {
  if (a.m <= 0 || a.n <= 0) return;
  const bool f32 = a.dtype == kF32;
  const int vec = f32 ? 4 : 8;
  bool vec_ok = (a.n % vec) == 0 && aligned16(a.out) && (a.ldo % vec) == 0;
  if (a.op != kOpZero) vec_ok = vec_ok && a.mode0 != kBcastImm && operand_vec_ok(a.in0, a.mode0, a.ld0, vec);
  if (a.op >= kOpAdd) vec_ok = vec_ok && operand_vec_ok(a.in1, a.mode1, a.ld1, vec);
  if (vec_ok) {
    const int grid = grid_for(a.m * (a.n / vec));
    if (f32) eltwise_vec_kernel<float, 4><<<grid, kThreads, 0, stream>>>(a);
    else eltwise_vec_kernel<uint16_t, 8><<<grid, kThreads, 0, stream>>>(a);
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
  const int grid = (int)(tiles < (int64_t)kNumSMs * 8 ? tiles : (int64_t)kNumSMs * 8);
  if (es == 4)
    transpose_kernel<uint32_t><<<grid, 256, 0, stream>>>(static_cast<const uint32_t *>(in),
                                                         static_cast<uint32_t *>(out), m, n, ldi, ldo);
  else
    transpose_kernel<uint16_t><<<grid, 256, 0, stream>>>(static_cast<const uint16_t *>(in),
                                                         static_cast<uint16_t *>(out), m, n, ldi, ldo);
  TPP_CUDA_CHECK(cudaGetLastError());
}

void launch_vnni2_pack(const void *in, void *out, int64_t m, int64_t n, int64_t ldi, int64_t ldo,
                       cudaStream_t stream) {
  if (m <= 0 || n <= 0) return;
  const uint16_t *src = static_cast<const uint16_t *>(in);
  uint16_t *dst = static_cast<uint16_t *>(out);
  const bool vec_ok = (n % 8) == 0 && (ldi % 8) == 0 && (ldo % 4) == 0 && aligned16(in) && aligned16(out);
  if (vec_ok)
    vnni2_pack_vec_kernel<<<grid_for((m / 2) * (n / 8)), kThreads, 0, stream>>>(src, dst, m, n, ldi, ldo);
  else
    vnni2_pack_scalar_kernel<<<grid_for(m * n), kThreads, 0, stream>>>(src, dst, m, n, ldi, ldo);
  TPP_CUDA_CHECK(cudaGetLastError());
}

void launch_vnni2_unpack(const void *in, void *out, int64_t m, int64_t n, int64_t ldi, int64_t ldo,
                         cudaStream_t stream) {
  if (m <= 0 || n <= 0) return;
  const uint16_t *src = static_cast<const uint16_t *>(in);
  uint16_t *dst = static_cast<uint16_t *>(out);
  const bool vec_ok = (n % 8) == 0 && (ldo % 8) == 0 && (ldi % 4) == 0 && aligned16(in) && aligned16(out);
  if (vec_ok)
    vnni2_unpack_vec_kernel<<<grid_for((m / 2) * (n / 8)), kThreads, 0, stream>>>(src, dst, m, n, ldi, ldo);
  else
    vnni2_unpack_scalar_kernel<<<grid_for(m * n), kThreads, 0, stream>>>(src, dst, m, n, ldi, ldo);
  TPP_CUDA_CHECK(cudaGetLastError());
}

} // namespace tpp
