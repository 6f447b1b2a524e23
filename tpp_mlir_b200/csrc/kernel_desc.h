// kernel_desc.h - the immutable descriptor a dispatch handle points to.
//
// A handle (the i64 the xsmm.*.dispatch ops return) is the address of one of
// these. libxsmm's equivalent is the JIT code registry entry returned by
// libxsmm_dispatch_* (runtime/Xsmm/XsmmRunnerUtils.cpp:131,169,201,351,446);
// like there, a handle lives for the life of the process and identical dispatch
// arguments return the identical handle.
#pragma once
#include <cstdint>

namespace tpp {

enum class OpClass : int32_t { Gemm = 1, Brgemm = 2, FusedBrgemm = 3, Unary = 4, Binary = 5, TileConfig = 6 };

// Which kernel family an invoke of this descriptor launches.
enum class KernelImpl : int32_t {
  None = 0,
  BrgemmTC = 1,     // tcgen05 / TMEM / TMA (bf16, TMA-compatible strides)
  BrgemmSimt = 2,   // generic FFMA kernel (f32, VNNI-B, odd strides)
  Eltwise = 3,      // unary identity/zero/relu, binary add/mul/sub/div
  Transpose = 4,
  Vnni2Pack = 5,
  Vnni2Unpack = 6,
  Noop = 7,
  Vnni4Pack = 8,
  Vnni4Unpack = 9,
};

constexpr uint32_t kDescMagic = 0x54505042u; // "TPPB"

struct KernelDesc {
  uint32_t magic = kDescMagic;
  OpClass op = OpClass::Gemm;
  KernelImpl impl = KernelImpl::None;
  int64_t dtype = 0;
  // gemm family
  int64_t m = 0, n = 0, k = 0, lda = 0, ldb = 0, ldc = 0, stride_a = 0, stride_b = 0;
  int64_t gemm_flags = 0;
  int64_t unary_flags = 0, unary_kind = 0, binary_flags = 0, binary_kind = 0;
  // eltwise family (m, n reused): unary ldi/ldo, binary ldi0/ldi1/ldo
  int64_t kind = 0, ldi = 0, ldi2 = 0, ldo = 0, flags = 0;
  // tcgen05 tile configuration chosen at dispatch
  int32_t block_n = 0;   // UMMA N (64/128/256)
  int32_t stages = 0;
  int32_t split_k = 1;   // cluster size along the reduction (DSMEM reduce)
  int32_t vnni_factor = 0;   // B operand [K/v][N][v] (gemm flag 2048): v = 2 or 4, the answer of libxsmm_cpuid_dot_pack_factor at dispatch; 0 = flat
  // VNNI-B descriptors: the same shape with a flat [K][N] B, run on the tcgen05 kernel after B has been
  // un-interleaved into a scratch buffer (nullptr when the shape is not tensor-core eligible)
  const KernelDesc *flat_twin = nullptr;
  char name[64] = {0};
};

} // namespace tpp
