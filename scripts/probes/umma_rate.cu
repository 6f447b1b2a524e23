// Probe: tensor-core time per tcgen05.mma (kind::f16, bf16 in, f32 acc) for small tile shapes, operands in shared
// memory with SWIZZLE_128B, issued back to back by one elected lane; reports clocks per instruction.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tpp_mlir_b200/csrc scripts/probes/umma_rate.cu -o gpurun_out/umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace tpp;

template <int M, int N, int A_MN, int B_MN>
__global__ void __launch_bounds__(128, 1) probe(unsigned long long *out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 160 * 1024;
  const uint32_t slot = bar + 8;
  volatile uint32_t *slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (slot - ptx::smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5;
  // zero the operand area (bf16 zeros: no NaN side effects)
  for (uint32_t i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x)
    asm volatile("st.shared.v4.u32 [%0], {%1,%1,%1,%1};" ::"r"(base + i * 16), "r"(0u));
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(slot, 256); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async();
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem = *slot_ptr;
  if (warp == 1) {
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(M, N, A_MN, B_MN);
    // A tiles at base (64 KiB area), B tiles at base + 64 KiB; 16 k-blocks, 4 k-steps each, cycled
    const uint64_t da0 = A_MN ? ptx::umma_smem_desc_sw128(base, 8192, 1024) : ptx::umma_smem_desc_sw128(base, 16, 1024);
    const uint64_t db0 = B_MN ? ptx::umma_smem_desc_sw128(base + 96 * 1024, 8192, 1024)
                              : ptx::umma_smem_desc_sw128(base + 96 * 1024, 16, 1024);
    const uint32_t a_blk = A_MN ? 8192 : (M * 128), b_blk = B_MN ? 8192 : (N * 128);   // bytes per 64-wide k-block
    const uint32_t a_step = A_MN ? 2048 : 32, b_step = B_MN ? 2048 : 32;
    unsigned long long t0 = 0, t1 = 0;
    for (int pass = 0; pass < 2; ++pass) {
      t0 = clock64();
      if (ptx::elect_one()) {
        for (int r = 0; r < reps; ++r) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              const uint64_t da = da0 + (uint64_t)((i * a_blk + kk * a_step) >> 4);
              const uint64_t db = db0 + (uint64_t)((i * b_blk + kk * b_step) >> 4);
              ptx::umma_bf16(tmem, da, db, idesc, 1u);
            }
          }
        }
        ptx::umma_commit(bar);
      }
      __syncwarp();
      ptx::mbar_wait(bar, pass & 1);
      t1 = clock64();
    }
    if ((threadIdx.x & 31) == 0) out[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after_sync(); ptx::tmem_dealloc(tmem, 256); }
}

template <int M, int N, int A_MN, int B_MN> void run(const char *name, int ctas) {
  unsigned long long *d;
  cudaMalloc(&d, 8 * 148);
  const int smem = 160 * 1024 + 1024 + 64;
  cudaFuncSetAttribute(probe<M, N, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 16;   // 256 MMAs
  probe<M, N, A_MN, B_MN><<<ctas, 128, smem>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, d, 8 * ctas, cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < ctas; ++i) s += (double)h[i];
  printf("%-34s ctas=%3d  %s  clk per MMA = %.1f  (smem operand bytes per MMA = %d)\n", name, ctas,
         e == cudaSuccess ? "ok" : cudaGetErrorString(e), s / ctas / (reps * 16), (M + N) * 32);
  cudaFree(d);
}

int main() {
  for (int ctas : {1, 128}) {
    run<64, 32, 1, 0>("M64 N32  A=MN-major B=K-major", ctas);
    run<64, 32, 0, 0>("M64 N32  A=K-major  B=K-major", ctas);
    run<64, 64, 1, 0>("M64 N64  A=MN-major B=K-major", ctas);
    run<64, 16, 1, 0>("M64 N16  A=MN-major B=K-major", ctas);
    run<128, 32, 1, 0>("M128 N32 A=MN-major B=K-major", ctas);
    run<128, 16, 1, 0>("M128 N16 A=MN-major B=K-major", ctas);
    run<128, 64, 0, 1>("M128 N64 A=K-major  B=MN-major", ctas);
    run<128, 32, 0, 1>("M128 N32 A=K-major  B=MN-major(64 wide)", ctas);
    run<128, 64, 0, 0>("M128 N64 A=K-major  B=K-major", ctas);
    run<128, 128, 0, 0>("M128 N128 A=K-major B=K-major", ctas);
    run<128, 256, 0, 0>("M128 N256 A=K-major B=K-major", ctas);
    run<64, 256, 0, 0>("M64 N256 A=K-major  B=K-major", ctas);
  }
  return 0;
}
