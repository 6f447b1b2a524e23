// mlp_chain.cu - first fused-chain design: L consecutive BRGEMM layers in ONE persistent launch, the
// per-layer tiling (128 x 64 tiles, 4-CTA split-K clusters, L2 exchange) with a grid-wide arrival counter between
// layers. Kept as the fallback for chain shapes the pass / pair kernels reject; also home of the kernel-independent
// chain tests (brgemm_chain_linked / brgemm_chain_supported).
#include "tc_common.cuh"
#include "tc_splitk.cuh"

namespace tpp {
using namespace tc;

namespace {

// ---- fused chain kernel: L consecutive BRGEMM layers in ONE persistent launch ----------------------------
// SURVEY.md section 8(f) item 2. The batch-256 MLP layer is latency-bound: per layer ~0.9 us of
// kernel hand-off plus ~4.4 us of kernel, most of it waiting. When a captured invoke sequence (xsmm_cuda_graph_*)
// contains layers whose C is the next layer's A, the runtime launches this kernel instead: same tiling as the
// stand-alone kernel (128 x 64 tiles, 4-CTA split-K clusters, L2 exchange), but
//   * the weight tiles of ALL layers (iters_per_cta x 8 KiB per layer) are fetched at kernel start and stay in
//     shared memory: no weight traffic on the critical path of layers 1..L-1;
//   * layers are separated by a grid-wide arrival counter instead of a kernel boundary (all CTAs are co-resident:
//     <= 148 CTAs, 1 per SM), so there is no launch hand-off and no per-layer setup;
//   * TMEM, barriers and the A stages are allocated once.
constexpr int CHAIN_IPC = 4;   // (batch x k-block) iterations per CTA and layer == A stages

struct ChainParams {
  CUtensorMap tmA[CHAIN_MAX_LAYERS], tmB[CHAIN_MAX_LAYERS];
  TcParams layer[CHAIN_MAX_LAYERS];
  unsigned int *grid_counter;   // monotonic arrival counter shared by all CTAs of this grid size
  int num_layers;
  int weights_early;            // no layer's weights / bias are written by in-flight kernels
  unsigned long long *trace;    // TPP_XSMM_TC_TRACE=1: clock stamps of layer 1 (nullptr in normal runs)
};

__device__ __forceinline__ void chain_stamp(const ChainParams &cp, int slot) {
  if (cp.trace) {
    const unsigned cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    cp.trace[(size_t)cta * TRACE_SLOTS + slot] = clock64();
  }
}

__global__ void __launch_bounds__(NUM_THREADS, 1) mlp_chain_kernel(const __grid_constant__ ChainParams cp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;                                     // CHAIN_IPC stages x 16 KiB
  const uint32_t smem_w = smem_base + CHAIN_IPC * A_STAGE_BYTES;         // L x CHAIN_IPC tiles x 8 KiB
  const uint32_t bar_base = smem_w + CHAIN_MAX_LAYERS * CHAIN_IPC * B_CHUNK_BYTES;
  const uint32_t a_full = bar_base;                                      // CHAIN_IPC
  const uint32_t w_full = bar_base + 8 * CHAIN_IPC;                      // CHAIN_MAX_LAYERS
  const uint32_t acc_bar = w_full + 8 * CHAIN_MAX_LAYERS;
  const uint32_t layer_bar = acc_bar + 8;
  const uint32_t tmem_slot = layer_bar + 8;
  uint8_t *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t m0 = blockIdx.y * BLOCK_M, n0 = blockIdx.x * 64;
  const uint32_t rank = blockIdx.z;                   // k-slice of this CTA (cluster = (1,1,4))
  const int L = cp.num_layers;
  const unsigned int G = gridDim.x * gridDim.y * gridDim.z;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < L; ++l) {
      ptx::prefetch_tensormap(&cp.tmA[l]);
      ptx::prefetch_tensormap(&cp.tmB[l]);
      ptx::mbar_init(w_full + 8 * l, 1);
    }
    for (int s = 0; s < CHAIN_IPC; ++s) ptx::mbar_init(a_full + 8 * s, 1);
    ptx::mbar_init(acc_bar, 1);
    ptx::mbar_init(layer_bar, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot_ptr;
  if (threadIdx.x == 0) chain_stamp(cp, 0);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  // weights of every layer: this CTA's k-slice (CHAIN_IPC k-blocks) x its 64 columns
  auto issue_weights = [&]() {
    for (int l = 0; l < L; ++l) {
      const TcParams &p = cp.layer[l];
      ptx::mbar_arrive_expect_tx(w_full + 8 * l, CHAIN_IPC * B_CHUNK_BYTES);
      for (int i = 0; i < CHAIN_IPC; ++i) {
        const int32_t it = (int32_t)rank * CHAIN_IPC + i;
        const int32_t b = it / p.k_iters, kb = it - b * p.k_iters;
        ptx::tma_load_3d(smem_w + (l * CHAIN_IPC + i) * B_CHUNK_BYTES, &cp.tmB[l], w_full + 8 * l, n0, kb * BLOCK_K, b);
      }
    }
  };
  if (warp == 0 && lane == 0 && cp.weights_early) issue_weights();
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (warp == 0 && lane == 0 && !cp.weights_early) issue_weights();

  for (int l = 0; l < L; ++l) {
    const TcParams &p = cp.layer[l];
    const uint32_t par = l & 1;
    if (warp == 0) {
      // ===== producer: this layer's A k-slice (previous layer's output once the whole grid has stored it) =====
      if (lane == 0) {
        if (l > 0) {
          ptx::mbar_wait(layer_bar, (l - 1) & 1);
          // Y(l-1) was written by other SMs with generic-proxy stores, fenced at gpu scope before the arrival
          // counter moved and acquired by this CTA's thread 64: it is in L2, which is where TMA reads from.
          if (l == 1) chain_stamp(cp, 1);
        }
        for (int i = 0; i < CHAIN_IPC; ++i) {
          const int32_t it = (int32_t)rank * CHAIN_IPC + i;
          const int32_t b = it / p.k_iters, kb = it - b * p.k_iters;
          ptx::mbar_arrive_expect_tx(a_full + 8 * i, A_STAGE_BYTES);
          ptx::tma_load_3d(smem_a + i * A_STAGE_BYTES, &cp.tmA[l], a_full + 8 * i, kb * BLOCK_K, m0, b);
        }
        if (l == 1) chain_stamp(cp, 2);
      }
      __syncwarp();
      ptx::cluster_arrive();   // the split-K exchange barrier of this layer (all threads of the cluster)
      ptx::cluster_wait();
    } else if (warp == 1) {
      // ===== MMA issuer =====
      if (lane == 0) {
        constexpr uint32_t idesc = ptx::umma_idesc_bf16(BLOCK_M, 64, 0, 1);
        ptx::mbar_wait(w_full + 8 * l, 0);
        for (int i = 0; i < CHAIN_IPC; ++i) {
          ptx::mbar_wait(a_full + 8 * i, par);
          ptx::tc_fence_after_sync();
          if (l == 1 && i == 0) chain_stamp(cp, 3);
          if (l == 1 && i == CHAIN_IPC - 1) chain_stamp(cp, 4);
          const uint32_t a_addr = smem_a + i * A_STAGE_BYTES;
          const uint32_t b_addr = smem_w + (l * CHAIN_IPC + i) * B_CHUNK_BYTES;
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            const uint64_t da = ptx::umma_smem_desc_sw128(a_addr + kk * (UMMA_K * 2), 16, 1024);
            const uint64_t db = ptx::umma_smem_desc_sw128(b_addr + kk * (UMMA_K * 128), B_CHUNK_BYTES, 1024);
            ptx::umma_bf16(tmem_acc, da, db, idesc, (i > 0 || kk > 0) ? 1u : 0u);
          }
        }
        ptx::umma_commit(acc_bar);
      }
      __syncwarp();
      ptx::cluster_arrive();
      ptx::cluster_wait();
    } else {
      // ===== epilogue: split-K exchange + fused bias/ReLU/store, then the grid-wide layer barrier =====
      const int q = warp & 3;
      ptx::mbar_wait(acc_bar, par);
      ptx::tc_fence_after_sync();
      if (l == 1 && threadIdx.x == 64) chain_stamp(cp, 7);
      splitk_epilogue_l2<16>(p, tmem_acc, q, lane, m0, n0, rank, true);
      if (l + 1 < L) {
        ptx::tc_fence_before_sync();                       // TMEM reads done before the next layer's MMAs overwrite it
        asm volatile("bar.sync 1, 128;" ::: "memory");     // all 128 epilogue threads have issued their Y(l) stores
        if (threadIdx.x == 64) {
          // one gpu-scope fence by the arriving thread: the CTA barrier above ordered the other threads' stores
          // before it (cumulativity), so they are visible device-wide before the counter moves
          __threadfence();
          const unsigned int old = atomicAdd(cp.grid_counter, 1u);
          const unsigned int target = (old / G + 1u) * G;
          unsigned int seen, spins = 0;
          do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(cp.grid_counter) : "memory");
            if (++spins > (1u << 22)) __trap();            // co-residency assumption broken: fail loudly, never hang
          } while (seen < target);
          if (l == 1) chain_stamp(cp, 11);
          if (l == 0) chain_stamp(cp, 5);
          ptx::mbar_arrive(layer_bar);                     // release the producer for layer l+1
        }
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (threadIdx.x == 0) chain_stamp(cp, 12);
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_acc, 64);
  }
}


} // namespace

// ---- fused chain launch -------------------------------------------------------------------------------------

// A layer a chain kernel could take: bf16 BRGEMM with beta_0 whose operands TMA can address (the flat descriptor, or a
// VNNI-2 B descriptor whose flat twin is tensor-core eligible), bias add (bcast_col) / ReLU or no epilogue.
bool brgemm_layer_chainable(const KernelDesc &d, const GemmArgs &g) {
  if (d.dtype != kBF16 || !(d.gemm_flags & 4)) return false;
  if (d.impl != KernelImpl::BrgemmTC && d.flat_twin == nullptr) return false;
  // VNNI-2: rewritten in shared memory by the chain kernels' converter warps; VNNI-4: through a flat copy (vnni_flat.cu)
  if ((d.gemm_flags & 2048) && d.vnni_factor != 2 && d.vnni_factor != 4) return false;
  if (g.batch < 1) return false;
  if (!aligned16(g.A) || !aligned16(g.B) || !aligned16(g.C) || (d.ldc % 8) != 0) return false;
  if (d.op == OpClass::FusedBrgemm && d.binary_kind != 0 && !(d.binary_kind == 1 && (d.binary_flags & 4))) return false;
  if (d.op == OpClass::FusedBrgemm && d.unary_kind != 0 && d.unary_kind != 5) return false;
  return true;
}

namespace {
// Address (in elements, relative to the operand base) of column `col` of a layer's OUTPUT row 0 ...
inline int64_t out_col_addr(const KernelDesc &d, const GemmArgs &g, int64_t col) {
  return (col / d.n) * g.c_step_k + (col % d.n);
}
// ... and of reduction index `kcol` of a layer's INPUT row 0 (batch element kcol / k, column kcol % k)
inline int64_t in_col_addr(const KernelDesc &d, int64_t kcol) { return (kcol / d.k) * d.stride_a + (kcol % d.k); }
}  // namespace

// Kernel-independent part: layers[0..L) form a chain - every layer a chainable bf16 BRGEMM on the same rows, layer l+1
// reads exactly layer l's output as its input (same base, same row structure, and every column of C(l) is the
// reduction index of A(l+1) that lives at the same address - true for the flat link "batch element b = columns
// [b k, b k + k)" and for the block-packed link "batch element b = output block b"), no weight /
// bias / output buffer is written inside the chain. Each chain kernel adds its own shape constraints on top
// (brgemm_chain_supported, chain_ft_supported, chain_pair_supported); a linked chain that no kernel takes is launched
// layer by layer.
bool brgemm_chain_linked(const KernelDesc *const *descs, const GemmArgs *args, int L) {
  static const bool off = [] { const char *e = getenv("TPP_XSMM_CHAIN"); return e && e[0] == '0'; }();
  if (off || L < 2 || L > CHAIN_MAX_LAYERS) return false;
  const KernelDesc &d0 = *descs[0];
  for (int l = 0; l < L; ++l) {
    const KernelDesc &d = *descs[l];
    const GemmArgs &g = args[l];
    if (!brgemm_layer_chainable(d, g) || d.m != d0.m || g.grid_n != args[0].grid_n) return false;
    if (((g.batch * d.k) % BLOCK_K) != 0) return false;
    if (l > 0) {
      const KernelDesc &p = *descs[l - 1];
      const GemmArgs &pg = args[l - 1];
      if (g.A != pg.C || d.lda != p.ldc) return false;
      if (g.grid_n > 1 && g.a_step != pg.c_step_n) return false;          // same row blocks
      const int64_t n_total = (int64_t)pg.grid_k * p.n;
      if (g.batch * d.k != n_total) return false;
      // column c of C(l-1) must be reduction index c of A(l): compare the two address maps block by block
      const int64_t step = p.n < d.k ? p.n : d.k;
      if ((p.n % step) != 0 || (d.k % step) != 0) return false;
      for (int64_t c = 0; c < n_total; c += step)
        if (out_col_addr(p, pg, c) != in_col_addr(d, c)) return false;
    }
  }
  return chain_operands_hazard_free(descs, args, L);
}

bool brgemm_chain_supported(const KernelDesc *const *descs, const GemmArgs *args, int L) {
  static const bool off = [] { const char *e = getenv("TPP_XSMM_CHAIN"); return e && e[0] == '0'; }();
  if (off || L < 2 || L > CHAIN_MAX_LAYERS) return false;
  const KernelDesc &d0 = *descs[0];
  const int64_t tiles = ((d0.n + 63) / 64) * ((d0.m + BLOCK_M - 1) / BLOCK_M);
  if (tiles * 4 > 128) return false;   // 4-CTA clusters: 33 fit at a time (measured); all CTAs must be co-resident
  for (int l = 0; l < L; ++l) {
    const KernelDesc &d = *descs[l];
    if (d.impl != KernelImpl::BrgemmTC || !(d.gemm_flags & 4) || d.m != d0.m || d.n != d0.n) return false;
    if (args[l].is_grid()) return false;   // the pass kernels address flat operands only
    const int64_t k_iters = (d.k + BLOCK_K - 1) / BLOCK_K;
    if ((d.k % BLOCK_K) != 0 || args[l].batch * k_iters != 4 * CHAIN_IPC) return false;
    if (!aligned16(args[l].A) || !aligned16(args[l].B) || !aligned16(args[l].C) || (d.ldc % 8) != 0) return false;
    if (d.op == OpClass::FusedBrgemm && d.binary_kind != 0 && !(d.binary_kind == 1 && (d.binary_flags & 4))) return false;
    if (l > 0) {
      // the chain link: A(l) is exactly C(l-1), viewed with the same leading dimension
      if (args[l].A != args[l - 1].C || d.lda != descs[l - 1]->ldc) return false;
      if (args[l].batch * d.k != descs[l - 1]->n) return false;
      if (args[l].batch > 1 && d.stride_a != d.k) return false;
    }
  }
  return chain_operands_hazard_free(descs, args, L);   // weights / bias must not be produced inside the chain
}


bool launch_brgemm_chain(const KernelDesc *const *descs, const GemmArgs *args, int L, cudaStream_t stream) {
  if (!brgemm_chain_supported(descs, args, L)) return false;
  {
    const int first = 0;
    if (launch_brgemm_chains_ft(descs, args, &first, &L, 1, stream) == 1) return true;
  }
  ChainParams cp;
  memset(&cp, 0, sizeof(cp));
  const KernelDesc &d0 = *descs[0];
  dim3 grid((unsigned)((d0.n + 63) / 64), (unsigned)((d0.m + BLOCK_M - 1) / BLOCK_M), 4);
  const int n_ctas = (int)(grid.x * grid.y * grid.z);
  // per-(thread, stream) exchange workspace + grid counters (same life cycle as the stand-alone kernel's workspace)
  float *ws = capture_owned_ws((size_t)148 * BLOCK_M * 64 * sizeof(float));
  unsigned int *counters = static_cast<unsigned int *>(capture_owned_zeroed(sizeof(unsigned int) * 256));
  cp.grid_counter = counters + n_ctas;   // one counter per grid size: always a multiple of G between launches
  cp.num_layers = L;
  cp.weights_early = 1;
  for (int l = 0; l < L; ++l) {
    const KernelDesc &d = *descs[l];
    const GemmArgs &g = args[l];
    const uint64_t nb = (uint64_t)g.batch;
    if (!encode_map(&cp.tmA[l], g.A, (uint64_t)d.k, (uint64_t)d.m, nb, (uint64_t)d.lda, (uint64_t)d.stride_a, BLOCK_K,
                    BLOCK_M) ||
        !encode_map(&cp.tmB[l], g.B, (uint64_t)d.n, (uint64_t)d.k, nb, (uint64_t)d.ldb, (uint64_t)d.stride_b, 64, BLOCK_K))
      return false;
    TcParams &p = cp.layer[l];
    p.C = g.C; p.D = g.D;
    p.m = d.m; p.n = d.n; p.ldc = d.ldc;
    p.k_iters = (int32_t)(d.k / BLOCK_K);
    p.total_iters = 4 * CHAIN_IPC;
    p.split_k = 4;
    p.beta0 = 1;
    p.bin_kind = (d.op == OpClass::FusedBrgemm && g.D) ? (int)d.binary_kind : 0;
    p.bin_mode = bin_mode_from_flags(d.binary_flags);
    p.relu = d.op == OpClass::FusedBrgemm && d.unary_kind == 5;
    p.c_vec_ok = 1;
    p.b_early = 0;
    p.flags = nullptr;
    p.ws = ws;
    p.trace = nullptr;
    if (!g.b_independent) cp.weights_early = 0;
  }
  constexpr int smem = CHAIN_IPC * A_STAGE_BYTES + CHAIN_MAX_LAYERS * CHAIN_IPC * B_CHUNK_BYTES +
                       (CHAIN_IPC + CHAIN_MAX_LAYERS + 2) * 8 + 16 + 1024;
  static std::once_flag once;
  std::call_once(once, [] {
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  });
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[3];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  attrs[1].id = cudaLaunchAttributeClusterDimension;
  attrs[1].val.clusterDim.x = 1;
  attrs[1].val.clusterDim.y = 1;
  attrs[1].val.clusterDim.z = 4;
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  // TPP_XSMM_TC_TRACE=3: the chain kernels stamp SM clocks of layer 1 into a buffer that is baked into the captured
  // graph; xsmm_cuda_debug_dump_trace() prints the averages of the last replay (no synchronisation here: this
  // function runs inside a stream capture)
  static const bool trace_on = [] { const char *e = getenv("TPP_XSMM_TC_TRACE"); return e && atoi(e) == 3; }();
  if (trace_on) {
    if (!g_trace_buf) {
      TPP_CUDA_CHECK(cudaMalloc(&g_trace_buf, sizeof(unsigned long long) * kTraceRing * kTraceRingCtas * TRACE_SLOTS));
      TPP_CUDA_CHECK(cudaMemsetAsync(g_trace_buf, 0, sizeof(unsigned long long) * kTraceRing * kTraceRingCtas * TRACE_SLOTS, stream));
    }
    cp.trace = g_trace_buf;
    if (L > 1) cp.layer[1].trace = g_trace_buf;   // splitk_epilogue_l2 stamps slots 8 (pushed) 9 (cluster) 10 (stored)
    g_chain_trace_ctas = n_ctas;
    g_chain_trace_layers = L;
  }
  if (!prepare_resident_launch(reinterpret_cast<const void *>(mlp_chain_kernel), &cfg, attrs)) return false;   // grid barrier per layer
  TPP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mlp_chain_kernel, cp));
  set_last_name("mlp_chain_bf16_%dlayers_128x64x64_splitk4", L);
  return true;
}


} // namespace tpp
