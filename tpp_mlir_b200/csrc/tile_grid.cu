// tile_grid.cu - a run of tile copies that walks a REGULAR grid (what a lowered tensor.pack / tensor.unpack emits: one
// unary identity TPP per tile, lib/TPP/Transforms/LowerPacksAndUnpacks.cpp:143-250) as one TMA-to-TMA copy kernel.
// Source and destination of the whole run are each ONE rank-4 tensor (4-byte unit in the tile row | tile row | inner tile
// index | outer tile index) with its own strides; a CTA streams boxes of several tiles through shared memory: TMA gathers
// a box from the source layout and TMA scatters the same box into the destination layout - no pointer table, no
// per-element address arithmetic, full lines on whichever side is contiguous. DESIGN.md 4.4.
#include "tc_common.cuh"

namespace tpp {
using namespace tc;

namespace {
constexpr int kNumSMs = 148;
constexpr int TG_STAGES = 6;
constexpr int TG_BOX_BYTES = 16 * 1024;
constexpr int TG_SMEM = TG_STAGES * TG_BOX_BYTES + TG_STAGES * 8 + 128;

struct TileGridParams {
  CUtensorMap src, dst;
  int32_t nb_m, nb_j, nb_i;   // boxes along the tile rows / the inner tile index / the outer tile index
  int32_t box_m, box_j;       // box extents (rows, inner tiles)
  uint32_t box_bytes;
};

__global__ void __launch_bounds__(32, 2) tile_grid_tma_kernel(const __grid_constant__ TileGridParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t bars = smem_base + TG_STAGES * TG_BOX_BYTES;
  if (threadIdx.x != 0) return;
  ptx::prefetch_tensormap(&p.src);
  ptx::prefetch_tensormap(&p.dst);
  for (int s = 0; s < TG_STAGES; ++s) ptx::mbar_init(bars + 8 * s, 1);
  ptx::fence_mbar_init();
  const int total = p.nb_m * p.nb_j * p.nb_i;
  const int n_my = (int)blockIdx.x < total ? (total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  auto coords = [&](int k, int32_t &c1, int32_t &c2, int32_t &c3) {   // my k-th box
    const int b = (int)blockIdx.x + k * (int)gridDim.x;
    const int im = b % p.nb_m, t = b / p.nb_m;
    c1 = im * p.box_m;
    c2 = (t % p.nb_j) * p.box_j;
    c3 = t / p.nb_j;
  };
  auto load = [&](int k) {
    const int s = k % TG_STAGES;
    int32_t c1, c2, c3;
    coords(k, c1, c2, c3);
    // a box that sticks out of the tensor is filled with zeros there and still counts box_bytes on the barrier
    ptx::mbar_arrive_expect_tx(bars + 8 * s, p.box_bytes);
    ptx::tma_load_4d(smem_base + s * TG_BOX_BYTES, &p.src, bars + 8 * s, 0, c1, c2, c3);
  };
  int issued = 0;
  for (; issued < TG_STAGES && issued < n_my; ++issued) load(issued);
  for (int k = 0; k < n_my; ++k) {
    const int s = k % TG_STAGES;
    ptx::mbar_wait(bars + 8 * s, (uint32_t)(k / TG_STAGES) & 1u);
    int32_t c1, c2, c3;
    coords(k, c1, c2, c3);
    ptx::tma_store_4d(&p.dst, smem_base + s * TG_BOX_BYTES, 0, c1, c2, c3);   // out-of-bounds parts are not written
    ptx::bulk_commit_group();
    if (k >= 1 && issued < n_my) {
      ptx::bulk_wait_group_read<1>();   // every store but the last has read its stage: box k - 1's stage is free
      load(issued);                     // issued == k - 1 + TG_STAGES: the same stage
      ++issued;
    }
  }
  ptx::bulk_wait_group<0>();
}
}  // namespace

// A run of `count` tile copies (m x n elements of `es` bytes, row pitches ldi / ldo) whose tile t = i * J + j reads
// in0 + i * in_outer + j * in_inner and writes out0 + i * out_outer + j * out_inner (byte steps). Returns false (and
// launches nothing) when TMA cannot express the run; the caller then uses the pointer-table kernel.
bool launch_tile_grid(const void *in0, void *out0, int64_t J, int64_t I, int64_t in_inner, int64_t in_outer, int64_t out_inner,
                      int64_t out_outer, int64_t m, int64_t n, int64_t ldi, int64_t ldo, int es, cudaStream_t stream) {
  const int64_t row_bytes = n * es, pitch_i = ldi * es, pitch_o = ldo * es;
  if ((row_bytes % 16) != 0 || row_bytes > 1024 || (pitch_i % 16) != 0 || (pitch_o % 16) != 0) return false;
  if (!aligned16(in0) || !aligned16(out0) || m < 1 || J < 1 || I < 1) return false;
  auto stride_ok = [](int64_t v, int64_t dim) { return dim == 1 || (v > 0 && (v % 16) == 0 && v < (1ll << 40)); };
  if (!stride_ok(pitch_i, m) || !stride_ok(pitch_o, m) || !stride_ok(in_inner, J) || !stride_ok(out_inner, J) ||
      !stride_ok(in_outer, I) || !stride_ok(out_outer, I))
    return false;
  if (m > (1ll << 31) || J > (1ll << 31) || I > (1ll << 31)) return false;
  // box: whole tile rows, as many rows / inner tiles as fit TG_BOX_BYTES
  const int64_t box_m = std::min<int64_t>(std::min<int64_t>(m, 256), std::max<int64_t>(1, TG_BOX_BYTES / row_bytes));
  const int64_t box_j = std::min<int64_t>(std::min<int64_t>(J, 256), std::max<int64_t>(1, TG_BOX_BYTES / (row_bytes * box_m)));
  TileGridParams p;
  const uint64_t dims[4] = {(uint64_t)(row_bytes / 4), (uint64_t)m, (uint64_t)J, (uint64_t)I};
  const uint32_t box[4] = {(uint32_t)(row_bytes / 4), (uint32_t)box_m, (uint32_t)box_j, 1};
  // a dimension of size 1 may carry any legal stride
  const uint64_t si[3] = {(uint64_t)(m > 1 ? pitch_i : 16), (uint64_t)(J > 1 ? in_inner : 16), (uint64_t)(I > 1 ? in_outer : 16)};
  const uint64_t so[3] = {(uint64_t)(m > 1 ? pitch_o : 16), (uint64_t)(J > 1 ? out_inner : 16), (uint64_t)(I > 1 ? out_outer : 16)};
  if (!encode_map_u32_4d(&p.src, in0, dims, si, box) || !encode_map_u32_4d(&p.dst, out0, dims, so, box)) return false;
  p.nb_m = (int32_t)((m + box_m - 1) / box_m);
  p.nb_j = (int32_t)((J + box_j - 1) / box_j);
  p.nb_i = (int32_t)I;
  p.box_m = (int32_t)box_m;
  p.box_j = (int32_t)box_j;
  p.box_bytes = (uint32_t)(row_bytes * box_m * box_j);
  const int64_t total = (int64_t)p.nb_m * p.nb_j * p.nb_i;
  if (total > (1ll << 30)) return false;
  static std::once_flag once;
  std::call_once(once, [] {
    TPP_CUDA_CHECK(cudaFuncSetAttribute(tile_grid_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TG_SMEM));
  });
  const int grid = (int)std::min<int64_t>(total, 2 * kNumSMs);
  tile_grid_tma_kernel<<<grid, 32, TG_SMEM, stream>>>(p);
  TPP_CUDA_CHECK(cudaGetLastError());
  return true;
}

}  // namespace tpp
