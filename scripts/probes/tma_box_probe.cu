// Probe: where does TMA put a 4-D box whose INNER extent (64 bytes) is smaller than the 128-byte swizzle span?
// Tensor = block-packed activations [8 row blocks][32 k blocks][32 rows][32 k] bf16, value = linear element index.
// Case 0: one box (32 k, 2 k-blocks, 32 rows, 4 row blocks), SWIZZLE_128B.
// Case 1: two boxes (32 k, 1, 32 rows, 4) written to base and base + 64 bytes, SWIZZLE_128B.
// Case 2: one box as case 0 with SWIZZLE_64B.   Case 3: box (32,1,32,4) SWIZZLE_64B.
// The kernel dumps 40 KiB of shared memory (pre-filled with 0xFFFF); the host prints, for the first rows, where each
// 16-byte chunk landed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I tpp_mlir_b200/csrc scripts/probes/tma_box_probe.cu -o scripts/probes/bin/tma_box_probe -lcuda
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace tpp;

constexpr int DUMP = 40 * 1024;

__device__ __forceinline__ void tma_load_4d_plain(uint32_t dst, const void *map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ CUtensorMap tm, uint16_t *out, int mode, uint32_t bytes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  const uint32_t bar = base + DUMP;
  for (int i = threadIdx.x; i < DUMP / 2; i += blockDim.x) reinterpret_cast<uint16_t *>(gen)[i] = 0xFFFF;
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  ptx::fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(bar, bytes);
    if (mode == 0) tma_load_4d_plain(base, &tm, bar, 0, 2, 0, 0);
    else { tma_load_4d_plain(base, &tm, bar, 0, 2, 0, 0); tma_load_4d_plain(base + 64, &tm, bar, 0, 3, 0, 0); }
  }
  ptx::mbar_wait(bar, 0);
  __syncthreads();
  for (int i = threadIdx.x; i < DUMP / 2; i += blockDim.x) out[i] = reinterpret_cast<uint16_t *>(gen)[i];
}

int main() {
  const int NI = 8, NB = 32, M = 32, K = 32;
  std::vector<uint16_t> h((size_t)NI * NB * M * K);
  // value encodes (i, b, rr, kk/8): i:3 bits | b:5 | rr:5 | chunk:2  (kk/8 = 16-byte chunk of the 64-byte row)
  for (int i = 0; i < NI; ++i) for (int b = 0; b < NB; ++b) for (int r = 0; r < M; ++r) for (int k = 0; k < K; ++k)
    h[(((size_t)i * NB + b) * M + r) * K + k] = (uint16_t)((i << 12) | (b << 7) | (r << 2) | (k >> 3));
  uint16_t *d, *o;
  cudaMalloc(&d, h.size() * 2); cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  cudaMalloc(&o, DUMP);
  void *sym = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                           const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                           CUtensorMapL2promotion, CUtensorMapFloatOOBfill)>(sym);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, DUMP + 2048);
  for (int cs = 0; cs < 4; ++cs) {
    if (cs == 1) continue;   // two half-line destinations: TMA rejects a 64-byte-aligned destination (misaligned address)
    CUtensorMap tm;
    cuuint64_t dims[4] = {K, NB, M, NI}, strides[3] = {(cuuint64_t)M * K * 2, (cuuint64_t)K * 2, (cuuint64_t)NB * M * K * 2};
    const bool two = (cs == 1 || cs == 3);
    cuuint32_t box[4] = {K, two ? 1u : 2u, M, 4}, es[4] = {1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     cs < 2 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("== case %d: encode rc=%d\n", cs, (int)r);
    if (r != CUDA_SUCCESS) continue;
    probe<<<1, 128, DUMP + 2048>>>(tm, o, cs == 1 ? 1 : 0, cs == 1 ? 2 * 8192u : (two ? 8192u : 16384u));
    cudaError_t e = cudaDeviceSynchronize();
    printf("   kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<uint16_t> s(DUMP / 2);
    cudaMemcpy(s.data(), o, DUMP, cudaMemcpyDeviceToHost);
    int last = -1;
    for (int i = 0; i < DUMP / 2; ++i) if (s[i] != 0xFFFF) last = i;
    printf("   last written byte offset: %d\n", last * 2 + 1);
    // print the first 12 128-byte lines: per 16-byte chunk the decoded (i,b,rr,chunk) or --
    for (int line = 0; line < 12; ++line) {
      printf("   line %2d:", line);
      for (int c = 0; c < 8; ++c) {
        const uint16_t v = s[(line * 128 + c * 16) / 2];
        if (v == 0xFFFF) printf("  ----------");
        else printf("  i%d b%d r%02d c%d", v >> 12, (v >> 7) & 31, (v >> 2) & 31, v & 3);
      }
      printf("\n");
    }
    // and lines 32..35 (next row block) 
    for (int line = 64; line < 68; ++line) {
      printf("   line %2d:", line);
      for (int c = 0; c < 8; ++c) {
        const uint16_t v = s[(line * 128 + c * 16) / 2];
        if (v == 0xFFFF) printf("  ----------");
        else printf("  i%d b%d r%02d c%d", v >> 12, (v >> 7) & 31, (v >> 2) & 31, v & 3);
      }
      printf("\n");
    }
  }
  return 0;
}
