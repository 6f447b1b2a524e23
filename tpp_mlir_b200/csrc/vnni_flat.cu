// vnni_flat.cu - flat [K][N] copies of VNNI-packed weights for the fused layer-chain kernels.
// The compiler packs bf16 weights as [k/v][n][v] (v = 2 on x86 with AVX512-BF16 / AMX, 4 with mlir-gen --vnni=4;
// lib/TPP/Transforms/Utils/VNNIUtils.cpp:75-78), block by block when the operands are block-packed. A chain kernel either
// rewrites the raw rows in shared memory (v = 2: converter warps, mlp_chain_pair.cu / mlp_chain_ft.cu) or reads a flat
// copy made by ONE launch of this kernel in front of it: every distinct weight buffer of the launch is un-interleaved and
// un-blocked into a scratch owned by the graph being captured. The copy is rebuilt by every replay, so weights the
// caller changed between two launches are seen. DESIGN.md 4.1c.
#include "tc_common.cuh"

namespace tpp {
using namespace tc;

namespace {
template <int V>
__device__ __forceinline__ void vnni_unit_to_rows(const uint16_t *__restrict__ src, uint16_t *__restrict__ dst, int64_t n_total) {
  // src: 8 columns x V k values, column-major inside the unit ([c][t]); dst: V rows of 8 columns, row pitch n_total
  uint4 in[V];                                   // 8 V elements = V 16-byte loads
#pragma unroll
  for (int i = 0; i < V; ++i) in[i] = __ldg(reinterpret_cast<const uint4 *>(src) + i);
  const uint16_t *e = reinterpret_cast<const uint16_t *>(in);   // e[c * V + t]
#pragma unroll
  for (int t = 0; t < V; ++t) {
    uint4 o;
    o.x = (uint32_t)e[0 * V + t] | ((uint32_t)e[1 * V + t] << 16);
    o.y = (uint32_t)e[2 * V + t] | ((uint32_t)e[3 * V + t] << 16);
    o.z = (uint32_t)e[4 * V + t] | ((uint32_t)e[5 * V + t] << 16);
    o.w = (uint32_t)e[6 * V + t] | ((uint32_t)e[7 * V + t] << 16);
    *reinterpret_cast<uint4 *>(dst + (int64_t)t * n_total) = o;
  }
}

template <int V>
__global__ void __launch_bounds__(256) vnni_weights_to_flat_kernel(const VnniFlatJob *__restrict__ jobs) {
  const VnniFlatJob L = jobs[blockIdx.y];
  const int64_t n_total = (int64_t)L.gk * L.n, k_total = (int64_t)L.nb * L.k;
  const int64_t c8s = n_total / 8, units = (k_total / V) * c8s;   // a unit: one k group x 8 columns
  const uint16_t *src0 = static_cast<const uint16_t *>(L.src);
  uint16_t *dst0 = static_cast<uint16_t *>(L.dst);
  for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
    const int64_t kg = u / c8s, col = (u - kg * c8s) * 8, k0 = kg * V;
    const int64_t j = col / L.n, nn = col - j * L.n, be = k0 / L.k, kk = k0 - be * L.k;
    vnni_unit_to_rows<V>(src0 + j * L.b_step + be * L.stride_b + ((kk / V) * L.ldb + nn) * V, dst0 + k0 * n_total + col, n_total);
  }
}
}  // namespace

bool vnni_flat_job_ok(const KernelDesc &d, const GemmArgs &g) {
  const int v = d.vnni_factor;
  if (v != 2 && v != 4) return false;
  if ((d.n % 8) != 0 || (d.k % v) != 0 || ((d.ldb * v) % 8) != 0) return false;   // 16-byte units on both sides
  if (g.batch > 1 && ((d.stride_b) % 8) != 0) return false;
  if (g.grid_k > 1 && (g.b_step % 8) != 0) return false;
  return aligned16(g.B);
}

VnniFlatJob vnni_flat_job(const KernelDesc &d, const GemmArgs &g, void *dst) {
  VnniFlatJob j;
  j.src = g.B;
  j.dst = dst;
  j.ldb = d.ldb;
  j.stride_b = g.batch > 1 ? d.stride_b : 0;
  j.b_step = g.grid_k > 1 ? g.b_step : 0;
  j.n = (int32_t)d.n;
  j.k = (int32_t)d.k;
  j.nb = (int32_t)g.batch;
  j.gk = g.grid_k;
  j.v = d.vnni_factor;
  return j;
}

// `count` jobs of one VNNI factor as ONE launch; the job table is copied to graph-owned device memory now (not captured)
void launch_vnni_weights_to_flat(const VnniFlatJob *jobs, int count, cudaStream_t stream) {
  if (count <= 0) return;
  const VnniFlatJob *table = static_cast<const VnniFlatJob *>(capture_owned_table(jobs, sizeof(VnniFlatJob) * (size_t)count));
  int64_t max_units = 0;
  for (int i = 0; i < count; ++i)
    max_units = std::max<int64_t>(max_units, (int64_t)jobs[i].nb * jobs[i].k / jobs[i].v * ((int64_t)jobs[i].gk * jobs[i].n / 8));
  // enough CTAs per job to cover it in a few sweeps, few enough that a launch with hundreds of jobs stays one wave deep
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((max_units + 255) / 256, count > 8 ? 16 : 128));
  for (int first = 0; first < count; first += 65535) {
    const unsigned gy = (unsigned)std::min(count - first, 65535);
    if (jobs[0].v == 4) vnni_weights_to_flat_kernel<4><<<dim3(gx, gy), 256, 0, stream>>>(table + first);
    else vnni_weights_to_flat_kernel<2><<<dim3(gx, gy), 256, 0, stream>>>(table + first);
    TPP_CUDA_CHECK(cudaGetLastError());
    note_extra_launch();
  }
}

}  // namespace tpp
