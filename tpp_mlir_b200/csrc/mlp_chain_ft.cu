// mlp_chain_ft.cu - feature-major pass kernels for launches that carry a FEW layer chains (1 .. 11): one CTA owns a
// (32 S batch rows) x (64 features) tile and computes it transposed (D^T = W^T X^T, tcgen05.mma M = 64), layer
// boundaries are per-batch-tile arrival counters, up to three chains are interleaved. DESIGN.md 4.1c.
#include "tc_common.cuh"

namespace tpp {
using namespace tc;

namespace {

// ---- feature-major chain kernel: no split-K, one 16-CTA software barrier per layer ---------------------------
// Second design of the fused chain, replacing the split-K clusters above for the MLP shape:
// the split-K chain spends more than half of a layer in its two synchronisations (cluster barrier around the f32
// partial exchange, grid barrier around the layer). Here one CTA owns a (32 batch rows) x (64 features) output
// tile over the FULL reduction and computes it transposed, D^T[64 features x 32 rows] = W^T x X^T with
// tcgen05.mma M = 64, N = 32 (A = the weight tile exactly as TMA delivers it, MN-major; B = the activation tile,
// K-major), so
//   * there is no partial-sum exchange at all: one accumulator (32 TMEM columns), one rounding, one store;
//   * a batch tile's n/64 CTAs only depend on each other (rows of the MLP are independent): the layer boundary is
//     an arrival counter per batch tile (8 independent groups of 16 CTAs), not a grid-wide barrier;
//   * TMEM lane = feature, so the bias is one scalar per thread and all four epilogue warps have work;
//   * operands move in few, large TMA boxes: the activation slice through a 4-D map (k-in-block, row, k-block,
//     batch) whose box covers 4 k-blocks x 32 rows = 16 KiB, the weights in 32 KiB boxes (the measured cost of one
//     TMA issue + barrier hand-off is ~150-600 clk, far more than the 100 clk of MMA work per k-block);
//   * the kernel runs a list of PASSES (one pass = one layer of one chain). The operands of pass p+1 stream into
//     the 16 weight / activation slots as pass p's MMAs retire them (tcgen05.commit per group of 4 slots), i.e.
//     under the MMAs, the epilogue and the barrier latency of pass p;
//   * several INDEPENDENT chains captured in one graph (the benchmark's rotating operand sets, or any batch of
//     forward passes on different buffers) become ONE launch whose pass list interleaves two chains (A.L0, B.L0,
//     A.L1, B.L1, ...): while chain A's layer output travels store -> fence -> counter -> poll (~2500 clk, most of
//     a layer when one chain runs alone), the tensor pipe works on chain B; there is no kernel boundary (1.5-1.8 us
//     of programmatic-launch hand-off) between forward passes, and the next chain's first-layer operands load
//     under the previous chain's last layer.
// Per pass and CTA the tensor pipe reads 192 KiB of operands from shared memory and TMA writes 192 KiB into it:
// at 128 B/clk that is ~3000 clk, the bound of this tiling (measured: tcgen05.mma time = operand bytes / 128 B/clk,
// scripts/probes/umma_rate.cu); L2 -> SM delivery of the same 192 KiB runs at ~53 B/clk per SM with all SMs pulling.
constexpr int FT_M = 64;                          // features per CTA  (UMMA M)
constexpr int FT_N = 32;                          // batch rows per CTA (UMMA N)
constexpr int FT_KB = 16;                         // k-block slots: (batch x k) reduction of at most 16 x 64
constexpr int FT_GROUP = 4;                       // k-blocks per TMA box / barrier
constexpr int FT_NG = FT_KB / FT_GROUP;
constexpr int FT_X_BYTES = FT_N * BLOCK_K * 2;    // 4 KiB
constexpr int FT_W_BYTES = BLOCK_K * FT_M * 2;    // 8 KiB
constexpr int FT_OUT_BYTES = FT_N * FT_M * 2;     // 4 KiB output staging (transposed back to row-major before it leaves the SM)
constexpr int FT_CTR_STRIDE = 32;                 // one 128-byte line per batch-tile counter
constexpr int FT_CTR_SLOT = 160 * FT_CTR_STRIDE;  // counters of chain slot s start at s * FT_CTR_SLOT
constexpr int FT_MAX_WAYS = 4;                    // chains interleaved in one launch (= counter slots)
constexpr int FT_MAX_PASSES = 64;

struct alignas(64) FtPass {
  CUtensorMap tmX, tmW;
  void *C;
  const void *D;
  int64_t ldc;
  int32_t k_iters;          // k-blocks per batch element
  int32_t groups;           // (batch x k-blocks) / FT_GROUP
  uint8_t has_bias, relu;
  uint8_t arrive;           // a later pass reads this pass's output: arrive on the slot's counter after storing
  uint8_t x_dep;            // X is the output of an earlier pass of the same chain slot: wait for wait_arrivals
  uint8_t slot;             // chain slot (< FT_MAX_WAYS) = which counter set this pass's chain uses
  uint8_t pad[3];
  uint32_t wait_arrivals;   // arrivals per CTA on the slot's counter (this launch) that must be visible before X loads
  // GEN instantiations only (layers that are grids of tile invokes on block-packed operands / have VNNI-2 weights): the
  // tile BRGEMM's m and n, the steps between output blocks, and per group of 4 k-blocks the box coordinates
  // (k-block within the batch element, batch element) of its first k-block
  int32_t m, n;
  int64_t c_step_n, c_step_k;
  int16_t grp_kb[4], grp_be[4];
  // the same for the weights, which may be addressed differently from the activations (a flat copy of VNNI-2 weights)
  int32_t w_n;
  int16_t w_kb[4], w_be[4];
};

struct FtParams {
  FtPass pass[FT_MAX_PASSES];
  unsigned int *counters;   // [slot][batch tile][FT_CTR_STRIDE]: monotonic, multiples of gridDim.x between launches
  int num_passes;
  int weights_early;        // no weight / bias is produced by in-flight kernels: fetch pass 0's before the PDL wait
  int x0_early;             // same for pass 0's activations
  int proxy_fence;
  int w_multicast;          // launched as (1,2,1) clusters: the two batch tiles of a cluster share each weight box
  // split-K-2 variant only: counter values are derived from a per-slot launch epoch instead of being read back
  unsigned int *epoch;      // [FT_MAX_WAYS] barriers completed per counter by earlier launches; [FT_MAX_WAYS] exit ticket
  unsigned int arrivals_total[FT_MAX_WAYS];   // barriers per counter this launch adds to each slot
  unsigned long long *trace;
};

__device__ __forceinline__ void ft_stamp(unsigned long long *trace, int slot) {
  if (trace) trace[(size_t)(blockIdx.x + gridDim.x * blockIdx.y) * FT_TRACE_SLOTS + slot] = clock64();
}
__device__ __forceinline__ void ft_stamp_pass(unsigned long long *trace, int p, int e) {
  if (trace && p < 9) trace[(size_t)(blockIdx.x + gridDim.x * blockIdx.y) * FT_TRACE_SLOTS + 8 + 6 * p + e] = clock64();
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int *p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// GEN: the layers are grids of tile BRGEMMs on block-packed operands (GemmArgs::grid_*; what the reference's default
// --tiles=32,32,32 emits) - X comes through a 5-D map (k in block | row in block | k-block in batch
// element | batch element | row block), W through a 4-D map (n in block | k in block | batch element | column block), the
// output tile is stored block by block. NARROW: 32-wide k blocks, whose 64-byte rows use SWIZZLE_64B sub-tiles (a k-block
// slot holds two [32 rows][32 k] sub-tiles). VNNI: VNNI-2 weights ([k/2][n][2]); TMA drops the raw rows of a group into
// its slots (row R of a k-block holds exactly the bytes of rows 2R, 2R+1 of the swizzled tile) and FT_CONV_WARPS converter
// warps rewrite them in place, the same scheme as the pair kernel's (mlp_chain_pair.cu). <false, false, false> is the
// flat kernel of rounds 1-2, unchanged.
constexpr int FT_CONV_WARPS = 8;
constexpr int FT_THREADS_VNNI = NUM_THREADS + 32 * FT_CONV_WARPS;
template <bool GEN, bool VNNI, bool NARROW>
__global__ void __launch_bounds__(VNNI ? FT_THREADS_VNNI : NUM_THREADS, 1) mlp_chain_ft_kernel(const __grid_constant__ FtParams cp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_x = smem_base;                                   // FT_KB x 4 KiB
  const uint32_t smem_w = smem_base + FT_KB * FT_X_BYTES;              // FT_KB x 8 KiB
  const uint32_t smem_out = smem_w + FT_KB * FT_W_BYTES;               // 4 KiB: the output tile, [32 rows][64 features]
  const uint32_t bar_base = smem_out + FT_OUT_BYTES;
  const uint32_t x_full = bar_base;                                    // [FT_NG]
  const uint32_t w_full = bar_base + 8 * FT_NG;                        // [FT_NG]
  const uint32_t w_empty = bar_base + 16 * FT_NG;                      // [FT_NG] group's X and W slots consumed
  const uint32_t acc_full = bar_base + 24 * FT_NG;                     // [2] accumulator (pass parity) complete
  const uint32_t acc_free = acc_full + 16;                             // [2] accumulator read out by the epilogue
  const uint32_t raw_full = acc_free + 16;                             // [FT_NG] (VNNI) the group's raw weight rows have landed
  const uint32_t tmem_slot = raw_full + 8 * FT_NG;
  uint8_t *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t n0 = blockIdx.x * FT_M;              // first feature of this CTA
  const int32_t m0 = blockIdx.y * FT_N;              // first batch row of this CTA
  const unsigned int G = gridDim.x;                  // CTAs per batch tile == arrivals per barrier
  unsigned int *counter0 = cp.counters + (size_t)blockIdx.y * FT_CTR_STRIDE;
  const int P = cp.num_passes;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&cp.pass[0].tmX);
    ptx::prefetch_tensormap(&cp.pass[0].tmW);
    for (int g = 0; g < FT_NG; ++g) {
      ptx::mbar_init(x_full + 8 * g, 1);
      ptx::mbar_init(w_full + 8 * g, VNNI ? FT_CONV_WARPS : 1);   // VNNI: one arrival per converter warp, no TMA bytes
      ptx::mbar_init(w_empty + 8 * g, cp.w_multicast ? 2 : 1);   // multicast: both CTAs of the cluster retire a group
      ptx::mbar_init(raw_full + 8 * g, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(acc_full + 8 * b, 1);
      ptx::mbar_init(acc_free + 8 * b, 4);           // one arrival per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 2 * FT_N);            // two 32-column accumulators, alternating by pass
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot_ptr;
  // weight multicast: the peer's barriers must exist before this CTA's first multicast box can complete on them
  const uint32_t crank = cp.w_multicast ? ptx::cluster_ctarank() : 0u;
  if (cp.w_multicast) ptx::cluster_sync();
  if (threadIdx.x == 0) ft_stamp(cp.trace, 0);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===== producer: the whole warp walks the (uniform) control flow, one elected lane issues =====
    // box coordinates (batch element, k-block) of group g's first slot, without integer division
    auto group_coords = [&](const FtPass &ps, int g, int32_t &b, int32_t &kb) {
      b = 0;
      kb = g * FT_GROUP;
      while (kb >= ps.k_iters) { kb -= ps.k_iters; ++b; }
    };
    auto issue_w = [&](int p, int g) {
      const FtPass &ps = cp.pass[p];
      if constexpr (GEN) {
        // my 64 features: column block n0 / n, column n0 % n in it (n == 32: two column blocks per box, VNNI only)
        const int32_t cb = n0 / ps.w_n, n_in = n0 - cb * ps.w_n;
        const int32_t kb = ps.w_kb[g], be = ps.w_be[g];
        if (ptx::elect_one()) {
          if constexpr (VNNI) {
            // raw [k/2][n][2] rows: (element of the row | column block | k pair | batch element), counted on raw_full
            ptx::mbar_arrive_expect_tx(raw_full + 8 * g, FT_GROUP * FT_W_BYTES);
            ptx::tma_load_4d(smem_w + g * (FT_GROUP * FT_W_BYTES), &ps.tmW, raw_full + 8 * g, 2 * n_in, cb, kb * (BLOCK_K / 2), be);
          } else {
            ptx::mbar_arrive_expect_tx(w_full + 8 * g, FT_GROUP * FT_W_BYTES);
            ptx::tma_load_4d(smem_w + g * (FT_GROUP * FT_W_BYTES), &ps.tmW, w_full + 8 * g, n_in, kb * BLOCK_K, be, cb);
          }
        }
        __syncwarp();
        return;
      }
      int32_t b, kb;
      group_coords(ps, g, b, kb);
      if (ptx::elect_one()) {
        // every CTA expects the whole box on its own barrier; with multicast only cluster rank (g & 1) fetches it,
        // and the box lands in both CTAs' slots (same offsets) and completes on both CTAs' barriers
        ptx::mbar_arrive_expect_tx(w_full + 8 * g, FT_GROUP * FT_W_BYTES);
        if (!cp.w_multicast)
          ptx::tma_load_3d(smem_w + g * (FT_GROUP * FT_W_BYTES), &ps.tmW, w_full + 8 * g, n0, kb * BLOCK_K, b);
        else if ((uint32_t)(g & 1) == crank)
          ptx::tma_load_3d_mc(smem_w + g * (FT_GROUP * FT_W_BYTES), &ps.tmW, w_full + 8 * g, n0, kb * BLOCK_K, b,
                              (uint16_t)0x3);
      }
      __syncwarp();
    };
    auto issue_x = [&](int p, int g) {
      const FtPass &ps = cp.pass[p];
      if constexpr (GEN) {
        const int32_t rb = m0 / ps.m, r_in = m0 - rb * ps.m;     // my 32 rows lie inside one row block
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(x_full + 8 * g, FT_GROUP * FT_X_BYTES);
          ptx::tma_load_5d(smem_x + g * (FT_GROUP * FT_X_BYTES), &ps.tmX, x_full + 8 * g, 0, r_in, ps.grp_kb[g], ps.grp_be[g], rb);
        }
        __syncwarp();
        return;
      }
      int32_t b, kb;
      group_coords(ps, g, b, kb);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(x_full + 8 * g, FT_GROUP * FT_X_BYTES);
        ptx::tma_load_4d(smem_x + g * (FT_GROUP * FT_X_BYTES), &ps.tmX, x_full + 8 * g, 0, m0, kb, b);
      }
      __syncwarp();
    };
    // pass 0's activations go out group by group with its weights when they are not produced by in-flight kernels - or
    // when the loads are issued behind the wait anyway
    const bool x0_early = !cp.pass[0].x_dep && (cp.weights_early ? cp.x0_early != 0 : true);
    auto early_loads = [&]() {
      for (int g = 0; g < cp.pass[0].groups; ++g) {  // group by group: the MMAs start on the first 48 KiB
        issue_w(0, g);
        if (x0_early) issue_x(0, g);
      }
      // the next passes' weights: this CTA's share of the feature tile's slice goes to L2 now
      if constexpr (!GEN) {
        for (int p = 1; p < P && p < 3; ++p) {
          const FtPass &ps = cp.pass[p];
          for (int g = (int)blockIdx.y; g < ps.groups; g += (int)gridDim.y) {
            int32_t b, kb;
            group_coords(ps, g, b, kb);
            if (ptx::elect_one()) ptx::tma_prefetch_3d(&ps.tmW, n0, kb * BLOCK_K, b);
            __syncwarp();
          }
        }
      }
    };
    if (cp.weights_early) early_loads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (lane == 0) ft_stamp(cp.trace, 1);
    if (!cp.weights_early) early_loads();
    // arrivals of this launch so far are < G on either counter (nobody passes a barrier without this CTA)
    unsigned int base[FT_MAX_WAYS];
#pragma unroll
    for (int sl = 0; sl < FT_MAX_WAYS; ++sl) base[sl] = (ld_acquire_gpu(counter0 + sl * FT_CTR_SLOT) / G) * G;
    for (int p = 0; p < P; ++p) {
      const FtPass &ps = cp.pass[p];
      if (p + 1 < P && lane == 0) {
        ptx::prefetch_tensormap(&cp.pass[p + 1].tmX);
        ptx::prefetch_tensormap(&cp.pass[p + 1].tmW);
      }
      const unsigned int *ctr = counter0 + (int)ps.slot * FT_CTR_SLOT;
      unsigned int target = ps.wait_arrivals * G;
#pragma unroll
      for (int sl = 0; sl < FT_MAX_WAYS; ++sl)
        if (sl == (int)ps.slot) target += base[sl];
      bool ready = !ps.x_dep;
      int x_next = (p == 0 && x0_early) ? ps.groups : 0;     // X groups issued so far
      for (int g = 0; g < ps.groups; ++g) {
        if (p > 0) ptx::mbar_wait(w_empty + 8 * g, (p - 1) & 1);   // slots of group g retired by pass p-1's MMAs
        if (p > 0) issue_w(p, g);                    // pass 0's weights were issued by early_loads()
        // relaxed (L2-coherent) polls: an acquire on every iteration costs a fence per poll. The data this flag
        // guards was fenced to L2 by its writers before they arrived, and it is only read by TMA (L2, never L1),
        // issued after the check - a control dependency the hardware does not speculate across.
        if (!ready) ready = (int)(ld_relaxed_gpu(ctr) - target) >= 0;   // one non-blocking look per group
        if (ready) {
          if (x_next == 0 && ps.x_dep) {
            if (cp.proxy_fence) asm volatile("fence.proxy.async;" ::: "memory");
            if (lane == 0) ft_stamp_pass(cp.trace, p, 0);
          }
          for (; x_next <= g; ++x_next) issue_x(p, x_next);
        }
      }
      if (!ready) {
        unsigned int spins = 0;
        while ((int)(ld_relaxed_gpu(ctr) - target) < 0) {
          if (++spins > (1u << 22)) __trap();        // co-residency assumption broken: fail loudly, never hang
        }
        if (cp.proxy_fence) asm volatile("fence.proxy.async;" ::: "memory");   // generic stores (other SMs) -> TMA reads
        if (lane == 0) ft_stamp_pass(cp.trace, p, 0);
      }
      for (; x_next < ps.groups; ++x_next) issue_x(p, x_next);
      if (lane == 0) ft_stamp_pass(cp.trace, p, 1);
    }
  } else if (warp == 1) {
    // ===== MMA issuer: uniform control flow, one elected lane issues =====
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(FT_M, FT_N, 1, 0);     // A (weights) MN-major, B (X) K-major
    // descriptors of slot 0 / k-step 0; every other (slot, k-step) is a constant added to the 14-bit address field
    const uint64_t da0 = ptx::umma_smem_desc_sw128(smem_w, FT_W_BYTES, 1024);
    // NARROW: a k-block slot of X is two [32 rows][32 k] SWIZZLE_64B sub-tiles (k steps 0,1 | 2,3), 8-row atoms of 512 bytes
    const uint64_t db0 = NARROW ? ptx::umma_smem_desc_sw64(smem_x, 16, 512) : ptx::umma_smem_desc_sw128(smem_x, 16, 1024);
    for (int p = 0; p < P; ++p) {
      const int NG = cp.pass[p].groups;
      const uint32_t par = p & 1;
      const uint32_t acc = tmem_acc + par * FT_N;
      if (p >= 2) {                                  // the epilogue of pass p-2 has read this accumulator out
        ptx::mbar_wait(acc_free + 8 * par, ((p >> 1) - 1) & 1);
        ptx::tc_fence_after_sync();
      }
#pragma unroll
      for (int g = 0; g < FT_NG; ++g) {
        if (g < NG) {
          ptx::mbar_wait(w_full + 8 * g, par);
          ptx::mbar_wait(x_full + 8 * g, par);
          ptx::tc_fence_after_sync();
          if (g == 0 && lane == 0) ft_stamp_pass(cp.trace, p, 2);
          if (ptx::elect_one()) {
#pragma unroll
            for (int j = 0; j < FT_GROUP; ++j) {
#pragma unroll
              for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
                const uint64_t da = da0 + (uint64_t)(((g * FT_GROUP + j) * FT_W_BYTES + kk * (UMMA_K * 128)) >> 4);
                const uint64_t db = db0 + (uint64_t)(((g * FT_GROUP + j) * FT_X_BYTES +
                                                      (NARROW ? (kk >> 1) * (FT_X_BYTES / 2) + (kk & 1) * (UMMA_K * 2) : kk * (UMMA_K * 2))) >> 4);
                ptx::umma_bf16(acc, da, db, idesc, (g > 0 || j > 0 || kk > 0) ? 1u : 0u);
              }
            }
            // the group's slots may be refilled with the next pass's tiles (multicast: tell both CTAs of the cluster)
            if (cp.w_multicast) ptx::umma_commit_mc(w_empty + 8 * g, (uint16_t)0x3);
            else ptx::umma_commit(w_empty + 8 * g);
            if (g == NG - 1) ptx::umma_commit(acc_full + 8 * par);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp < 6) {
    // ===== epilogue: TMEM lanes 32q + (0..15) hold features 16q + (0..15); columns = the 32 batch rows =====
    const int q = warp & 3;
    const int f = 16 * q + (lane & 15);
    const bool active = lane < 16;
    const int ep_tid0 = 64;                          // first epilogue thread: the one that arrives for the CTA
    auto load_bias = [&](int p) -> uint16_t {
      return (p < P && cp.pass[p].has_bias) ? __ldg(static_cast<const uint16_t *>(cp.pass[p].D) + n0 + f) : (uint16_t)0;
    };
    // this thread's bias of the first pass, requested before the wait when the parameters are not produced in flight
    uint16_t bias_next = 0;
    if (cp.weights_early) bias_next = load_bias(0);
    // everything before this point only read memory; no store may precede the previous kernel's completion
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (!cp.weights_early) bias_next = load_bias(0);
    for (int p = 0; p < P; ++p) {
      const FtPass &ps = cp.pass[p];
      const float bias = bf16_bits_to_f32(bias_next);
      bias_next = load_bias(p + 1);                  // in flight while this pass's accumulator completes
      const uint32_t par = p & 1;
      ptx::mbar_wait(acc_full + 8 * par, (p >> 1) & 1);
      ptx::tc_fence_after_sync();
      if (threadIdx.x == ep_tid0) ft_stamp_pass(cp.trace, p, 3);
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + par * FT_N, r);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_free + 8 * par);
      // TMEM lane = feature, column = batch row: the tile is transposed back through shared memory so that it leaves
      // the SM as 16-byte pieces of row-major rows (2 store instructions per thread instead of 32 two-byte stores whose
      // acknowledgements the release fence below then had to wait for: stored -> arrived was 1100-1500 clk)
      if (active) {
        const uint32_t dst = smem_out + (uint32_t)f * 2u;
#pragma unroll
        for (int j = 0; j < FT_N; ++j) {
          const float v = __uint_as_float(r[j]) + bias;
          const uint16_t h = f32_to_bf16_bits(ps.relu ? relu_f32(v) : v);
          asm volatile("st.shared.u16 [%0], %1;" ::"r"(dst + (uint32_t)j * (FT_M * 2)), "h"(h) : "memory");
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");       // the whole tile is staged
      {
        const int et = (int)threadIdx.x - ep_tid0;          // 0 .. 127
        uint16_t *crow = static_cast<uint16_t *>(ps.C) + (int64_t)m0 * ps.ldc + n0;
        if constexpr (GEN) {
          // block-packed output: row block m0 / m, this thread's 8 features (the same for both of its pieces) lie in
          // column block col / n
          const int32_t rb = m0 / ps.m, r_in = m0 - rb * ps.m;
          const int32_t col = n0 + (et & 7) * 8, cb = col / ps.n, n_in = col - cb * ps.n;
          crow = static_cast<uint16_t *>(ps.C) + (int64_t)rb * ps.c_step_n + (int64_t)cb * ps.c_step_k + (int64_t)r_in * ps.ldc +
                 n_in - (et & 7) * 8;
        }
#pragma unroll
        for (int c = et; c < FT_N * FT_M * 2 / 16; c += 128) {
          const int row = c >> 3, col16 = c & 7;
          uint4 v;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                       : "r"(smem_out + (uint32_t)c * 16u));
          *reinterpret_cast<uint4 *>(crow + (int64_t)row * ps.ldc + col16 * 8) = v;
        }
      }
      if (threadIdx.x == ep_tid0) ft_stamp_pass(cp.trace, p, 4);
      if (ps.arrive) {
        asm volatile("bar.sync 1, 128;" ::: "memory");     // all epilogue threads have issued their stores
        if (threadIdx.x == ep_tid0) {
          // one gpu-scope release by the arriving thread; the CTA barrier ordered the other threads' stores before it
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(counter0 + (int)ps.slot * FT_CTR_SLOT) : "memory");
          ft_stamp_pass(cp.trace, p, 5);
        }
      } else {
        asm volatile("bar.sync 1, 128;" ::: "memory");     // the staging buffer is free for the next pass
      }
    }
  } else if (VNNI) {
    // ===== VNNI-2 weight converters: raw [k/2][n][2] rows of a group's four k-block slots -> swizzled MN-major tiles, in
    // place. A slot holds 32 k-pair rows of 256 bytes (64 features x 2 k); row R is exactly the bytes of the tile's rows
    // 2R and 2R + 1 (128 bytes each). One warp-wide 16-byte load covers two raw rows; lanes 2p, 2p + 1 hold features
    // 8p .. 8p + 3 / 8p + 4 .. 8p + 7 (both k of the pair), swap halves with one shuffle, the even lane writes the chunk
    // of the even k row, the odd lane that of the odd k row (chunk index XOR row & 7: what TMA's SWIZZLE_128B would have
    // produced for flat weights). Only shared memory is touched: the converters may run ahead of the PDL wait. =====
    const int cw = warp - 6;                          // 0 .. FT_CONV_WARPS - 1
    // (which 8-feature group a lane pair takes: within a quarter-warp the four pairs take groups {0,5,2,7} / {4,1,6,3}, so
    // that the quarter's eight 16-byte loads AND its eight 16-byte stores - two tile rows whose chunk positions differ
    // in bit 0 only - each cover all 32 banks once)
    const int row_sub = lane >> 4, half = lane & 1, g8 = vnni_group_of_lane(lane);
    for (int p = 0; p < P; ++p) {
      const int NG = cp.pass[p].groups;
      for (int g = 0; g < NG; ++g) {
        if (lane == 0) ptx::mbar_wait(raw_full + 8 * g, p & 1);   // one poller per warp
        __syncwarp();
        uint4 v[2 * FT_GROUP];
#pragma unroll
        for (int u = 0; u < 2 * FT_GROUP; ++u) {
          const uint32_t R = (uint32_t)((u & 1) * 16 + cw * 2 + row_sub);     // raw row = k pair of the k-block
          const uint32_t src = smem_w + (uint32_t)(g * FT_GROUP + (u >> 1)) * FT_W_BYTES + R * 256u + (uint32_t)(2 * g8 + half) * 16u;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "r"(src));
        }
        __syncwarp();                                 // every lane has read its rows before any lane overwrites them
#pragma unroll
        for (int u = 0; u < 2 * FT_GROUP; ++u) {
          const uint32_t lo0 = __byte_perm(v[u].x, v[u].y, 0x5410), lo1 = __byte_perm(v[u].z, v[u].w, 0x5410);
          const uint32_t hi0 = __byte_perm(v[u].x, v[u].y, 0x7632), hi1 = __byte_perm(v[u].z, v[u].w, 0x7632);
          const uint32_t r0 = __shfl_xor_sync(0xffffffffu, half ? lo0 : hi0, 1);
          const uint32_t r1 = __shfl_xor_sync(0xffffffffu, half ? lo1 : hi1, 1);
          const uint32_t o0 = half ? r0 : lo0, o1 = half ? r1 : lo1, o2 = half ? hi0 : r0, o3 = half ? hi1 : r1;
          const uint32_t krow = 2u * (uint32_t)((u & 1) * 16 + cw * 2 + row_sub) + (uint32_t)half;   // k row of the 64 x 64 tile
          const uint32_t base = smem_w + (uint32_t)(g * FT_GROUP + (u >> 1)) * FT_W_BYTES;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                       ::"r"(base + krow * 128u + (((uint32_t)g8 ^ (krow & 7u)) << 4)), "r"(o0), "r"(o1), "r"(o2), "r"(o3)
                       : "memory");
        }
        ptx::fence_proxy_async();                     // my shared-memory writes -> the async proxy (the MMAs)
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(w_full + 8 * g);
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (cp.w_multicast) ptx::cluster_sync();           // the peer may still signal this CTA's barriers until it is done too
  if (threadIdx.x == 0) ft_stamp(cp.trace, 2);
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_acc, 2 * FT_N);
  }
}

// ---- split-K variants (S = 2, 4) of the pass kernel -----------------------------------------------------------------
// A pass of the kernel above is bound by the bytes one SM receives (192 KiB at ~50 B/clk); the tensor pipe and even
// shared memory have slack. Here a cluster of S CTAs shares a (64 features) x (32 S batch rows) tile and splits the
// reduction S ways: per pass a CTA receives 128/S KiB of weights + 64 KiB of activations (S = 2: 128 KiB, S = 4:
// 96 KiB) and the S partial accumulators meet through distributed shared memory: each CTA owns 32 of the rows and
// pushes the other rows of its partial (8 KiB per peer) into the peers' receive buffers with st.async, whose bytes
// complete_tx on the RECEIVER's mbarrier - no global-memory round trip, no cluster-wide barrier, no release/acquire
// round trip (a release.cluster arrive after plain st.shared::cluster stores cost ~3000 clk per pass). Everything else
// (pass list, interleaved chains, slot retirement by tcgen05.commit, arrival counters) is unchanged; a consumer CTA
// (feature tile, batch tile, k-slice z) waits for the 16 CTAs that produce its slice of the features.
template <int S> struct FS {
  static constexpr int N = 32 * S;                  // batch rows per tile (UMMA N); a CTA stores 32 of them
  static constexpr int KB = FT_KB / S;              // k-block slots per CTA
  static constexpr int NG = 4;                      // groups (TMA boxes / barriers) per pass
  static constexpr int GROUP = KB / NG;             // k-blocks per group
  static constexpr int X_BYTES = N * BLOCK_K * 2;   // one k-block of activations
  static constexpr int RECV_BYTES = FT_M * 32 * 4;  // one peer's partial for my 32 rows: 64 features x 32 f32
  static constexpr int SMEM = KB * (X_BYTES + FT_W_BYTES) + 2 * (S - 1) * RECV_BYTES + (3 * NG + 8) * 8 + 16 + 1024;
};
constexpr int F2_THREADS = 352;                   // producer, MMA issuer, 4 finisher warps, 4 sender warps, arriver

template <int S>
__global__ void __launch_bounds__(F2_THREADS, 1) mlp_chain_fts_kernel(const __grid_constant__ FtParams cp) {
  constexpr int F2_N = FS<S>::N, F2_KB = FS<S>::KB, F2_NG = FS<S>::NG, F2_GROUP = FS<S>::GROUP;
  constexpr int F2_X_BYTES = FS<S>::X_BYTES, F2_RECV_BYTES = FS<S>::RECV_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_x = smem_base;                                   // F2_KB x 8 KiB
  const uint32_t smem_w = smem_base + F2_KB * F2_X_BYTES;              // F2_KB x 8 KiB
  const uint32_t smem_recv = smem_w + F2_KB * FT_W_BYTES;              // [pass parity][sender rank slot] x 8 KiB
  const uint32_t bar_base = smem_recv + 2 * (S - 1) * F2_RECV_BYTES;
  const uint32_t x_full = bar_base;                                    // [F2_NG]
  const uint32_t w_full = bar_base + 8 * F2_NG;
  const uint32_t w_empty = bar_base + 16 * F2_NG;
  const uint32_t acc_full = bar_base + 24 * F2_NG;                     // [2]
  const uint32_t acc_free = acc_full + 16;                             // [2]
  const uint32_t xchg_full = acc_free + 16;                            // [2] the peer's partial has landed in recv[parity]
  const uint32_t recv_free = xchg_full + 16;                           // [2] (local) my finishers are done with recv[parity]
  const uint32_t tmem_slot = recv_free + 16;
  uint8_t *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t n0 = blockIdx.x * FT_M;              // first feature of this CTA
  const int32_t m0 = blockIdx.y * F2_N;              // first batch row of the pair's tile
  const uint32_t z = blockIdx.z;                     // k-slice of this CTA == its rank in the (1,1,S) cluster
  const unsigned int G = gridDim.x;                  // arrivals per barrier: gridDim.x/S feature tiles x S k-slices
  // the counter this CTA waits on: its batch tile, ITS k-slice of the next layer's reduction
  unsigned int *wait_ctr0 = cp.counters + (size_t)(blockIdx.y * S + z) * FT_CTR_STRIDE;
  // the counter this CTA arrives on: its batch tile, the k-slice its features belong to
  unsigned int *arrive_ctr0 = cp.counters + (size_t)(blockIdx.y * S + blockIdx.x / (gridDim.x / S)) * FT_CTR_STRIDE;
  const int P = cp.num_passes;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&cp.pass[0].tmX);
    ptx::prefetch_tensormap(&cp.pass[0].tmW);
    for (int g = 0; g < F2_NG; ++g) {
      ptx::mbar_init(x_full + 8 * g, 1);
      ptx::mbar_init(w_full + 8 * g, 1);
      ptx::mbar_init(w_empty + 8 * g, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(acc_full + 8 * b, 1);
      ptx::mbar_init(acc_free + 8 * b, 8);           // one arrival per finisher and per sender warp
      ptx::mbar_init(xchg_full + 8 * b, 1);          // one expect_tx arrival (mine); the peers' st.async bytes complete it
      ptx::mbar_init(recv_free + 8 * b, 4);          // one arrival per finisher warp
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, 2 * F2_N);            // two 64-column accumulators, alternating by pass
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot_ptr;
  ptx::cluster_sync();                               // the peer's barriers exist before anything is pushed to it
  if (threadIdx.x == 0) ft_stamp(cp.trace, 0);

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===== producer =====
    // box coordinates (batch element, k-block) of group g's first k-block: global k-block index z * 8 + 2 g
    auto group_coords = [&](const FtPass &ps, int g, int32_t &b, int32_t &kb) {
      b = 0;
      kb = (int32_t)z * F2_KB + g * F2_GROUP;
      while (kb >= ps.k_iters) { kb -= ps.k_iters; ++b; }
    };
    auto issue_w = [&](int p, int g) {
      const FtPass &ps = cp.pass[p];
      int32_t b, kb;
      group_coords(ps, g, b, kb);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(w_full + 8 * g, F2_GROUP * FT_W_BYTES);
        ptx::tma_load_3d(smem_w + g * (F2_GROUP * FT_W_BYTES), &ps.tmW, w_full + 8 * g, n0, kb * BLOCK_K, b);
      }
      __syncwarp();
    };
    auto issue_x = [&](int p, int g) {
      const FtPass &ps = cp.pass[p];
      int32_t b, kb;
      group_coords(ps, g, b, kb);
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(x_full + 8 * g, F2_GROUP * F2_X_BYTES);
        ptx::tma_load_4d(smem_x + g * (F2_GROUP * F2_X_BYTES), &ps.tmX, x_full + 8 * g, 0, m0, kb, b);
      }
      __syncwarp();
    };
    const bool x0_early = cp.weights_early && cp.x0_early && !cp.pass[0].x_dep;
    auto early_loads = [&]() {
      for (int g = 0; g < F2_NG; ++g) {
        issue_w(0, g);
        if (x0_early) issue_x(0, g);
      }
    };
    if (cp.weights_early) early_loads();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (lane == 0) ft_stamp(cp.trace, 1);
    if (!cp.weights_early) early_loads();
    unsigned int base[FT_MAX_WAYS];
#pragma unroll
    for (int sl = 0; sl < FT_MAX_WAYS; ++sl) base[sl] = ld_acquire_gpu(cp.epoch + sl) * G;
    for (int p = 0; p < P; ++p) {
      const FtPass &ps = cp.pass[p];
      if (p + 1 < P && lane == 0) {
        ptx::prefetch_tensormap(&cp.pass[p + 1].tmX);
        ptx::prefetch_tensormap(&cp.pass[p + 1].tmW);
      }
      const unsigned int *ctr = wait_ctr0 + (int)ps.slot * FT_CTR_SLOT;
      unsigned int target = ps.wait_arrivals * G;
#pragma unroll
      for (int sl = 0; sl < FT_MAX_WAYS; ++sl)
        if (sl == (int)ps.slot) target += base[sl];
      bool ready = !ps.x_dep;
      int x_next = (p == 0 && x0_early) ? F2_NG : 0;
      for (int g = 0; g < F2_NG; ++g) {
        if (p > 0) ptx::mbar_wait(w_empty + 8 * g, (p - 1) & 1);
        if (p > 0) issue_w(p, g);
        if (!ready) ready = (int)(ld_relaxed_gpu(ctr) - target) >= 0;
        if (ready) {
          if (x_next == 0 && ps.x_dep) {
            if (cp.proxy_fence) asm volatile("fence.proxy.async;" ::: "memory");
          }
          for (; x_next <= g; ++x_next) issue_x(p, x_next);
        }
      }
      if (!ready) {
        unsigned int spins = 0;
        while ((int)(ld_relaxed_gpu(ctr) - target) < 0) {
          if (++spins > (1u << 22)) __trap();        // co-residency assumption broken: fail loudly, never hang
        }
        if (cp.proxy_fence) asm volatile("fence.proxy.async;" ::: "memory");
      }
      for (; x_next < F2_NG; ++x_next) issue_x(p, x_next);
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(FT_M, F2_N, 1, 0);     // A (weights) MN-major, B (X) K-major
    const uint64_t da0 = ptx::umma_smem_desc_sw128(smem_w, FT_W_BYTES, 1024);
    const uint64_t db0 = ptx::umma_smem_desc_sw128(smem_x, 16, 1024);
    for (int p = 0; p < P; ++p) {
      const uint32_t par = p & 1;
      const uint32_t acc = tmem_acc + par * F2_N;
      if (p >= 2) {
        ptx::mbar_wait(acc_free + 8 * par, ((p >> 1) - 1) & 1);
        ptx::tc_fence_after_sync();
      }
#pragma unroll
      for (int g = 0; g < F2_NG; ++g) {
        ptx::mbar_wait(w_full + 8 * g, par);
        ptx::mbar_wait(x_full + 8 * g, par);
        ptx::tc_fence_after_sync();
        if (ptx::elect_one()) {
#pragma unroll
          for (int j = 0; j < F2_GROUP; ++j) {
#pragma unroll
            for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
              const uint64_t da = da0 + (uint64_t)(((g * F2_GROUP + j) * FT_W_BYTES + kk * (UMMA_K * 128)) >> 4);
              const uint64_t db = db0 + (uint64_t)(((g * F2_GROUP + j) * F2_X_BYTES + kk * (UMMA_K * 2)) >> 4);
              ptx::umma_bf16(acc, da, db, idesc, (g > 0 || j > 0 || kk > 0) ? 1u : 0u);
            }
          }
          ptx::umma_commit(w_empty + 8 * g);
          if (g == F2_NG - 1) ptx::umma_commit(acc_full + 8 * par);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogue: lanes 0..15 of quarter q hold features 16q + lane; columns = the tile's 64 batch rows.
    // This CTA finishes rows [32z, 32z+32). SENDER warps (6..9) push the other 32 columns of the partial into the
    // peer's receive buffer and signal it; FINISHER warps (2..5) add the peer's partial to their own 32 columns, apply
    // bias / ReLU, round once and store. Two warp sets, so that waiting for the peer never delays what the peer waits for.
    const int q = warp & 3;
    const int f = 16 * q + (lane & 15);
    const bool active = lane < 16;
    if (warp == 10) {
      // ===== arriver: publishes a pass's output for the finishers (they only bar.arrive), so the ~1100-clk gpu-scope
      // fence is off their critical path =====
      asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int p = 0; p < P; ++p) {
        const FtPass &ps = cp.pass[p];
        if (!ps.arrive) continue;
        asm volatile("bar.sync 1, 160;" ::: "memory");     // the 128 finisher threads have issued this pass's stores
        if (lane == 0) {
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(arrive_ctr0 + (int)ps.slot * FT_CTR_SLOT) : "memory");
          ft_stamp_pass(cp.trace, p, 5);
        }
        __syncwarp();
      }
    } else if (warp >= 6) {
      for (int p = 0; p < P; ++p) {
        const uint32_t par = p & 1;
        ptx::mbar_wait(acc_full + 8 * par, (p >> 1) & 1);
        ptx::tc_fence_after_sync();
        uint32_t oth[S - 1][32];
#pragma unroll
        for (int r = 1; r < S; ++r)                   // the 32 columns (rows of the tile) owned by cluster rank z ^ r
          ptx::tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + par * F2_N + 32 * (z ^ (uint32_t)r),
                             oth[r - 1]);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(acc_free + 8 * par);
        if (threadIdx.x == 192) ft_stamp_pass(cp.trace, p, 1);
        // Overwriting the peer's recv[par] (last used by pass p-2) needs no signal from the peer: this push waits
        // until MY finishers have consumed pass p-1, i.e. received the peer's push of pass p-1, which the peer only
        // sent after ITS finishers had consumed pass p-2 (same rule on its side). All local, no cross-SM release.
        if (p >= 1) ptx::mbar_wait(recv_free + 8 * ((p - 1) & 1), ((p - 1) >> 1) & 1);
        if (active) {
          // st.async: every 16-byte store carries its own completion (complete_tx on the peer's barrier); a
          // release.cluster arrive after plain st.shared::cluster stores cost ~3000 clk per pass here
#pragma unroll
          for (int r = 1; r < S; ++r) {
            const uint32_t peer = z ^ (uint32_t)r;     // the peer files my partial under slot r - 1 (it sees me as peer ^ r)
            const uint32_t remote = ptx::mapa(smem_recv + (par * (S - 1) + (r - 1)) * F2_RECV_BYTES + (uint32_t)f * 128u, peer);
            const uint32_t remote_bar = ptx::mapa(xchg_full + 8 * par, peer);
#pragma unroll
            for (int j = 0; j < 8; ++j)                // 16-byte chunks XOR-swizzled by the feature: no bank conflicts
              ptx::st_async_v4(remote + (uint32_t)((j ^ (f & 7)) << 4), remote_bar, oth[r - 1][4 * j],
                               oth[r - 1][4 * j + 1], oth[r - 1][4 * j + 2], oth[r - 1][4 * j + 3]);
          }
        }
        if (threadIdx.x == 192) ft_stamp_pass(cp.trace, p, 2);
      }
    } else {
      const int ep_tid0 = 64;
      auto load_bias = [&](int p) -> uint16_t {
        return (p < P && cp.pass[p].has_bias) ? __ldg(static_cast<const uint16_t *>(cp.pass[p].D) + n0 + f) : (uint16_t)0;
      };
      uint16_t bias_next = 0;
      if (cp.weights_early) bias_next = load_bias(0);
      asm volatile("griddepcontrol.wait;" ::: "memory");   // no store before the previous kernel has completed
      if (!cp.weights_early) bias_next = load_bias(0);
      for (int p = 0; p < P; ++p) {
        const FtPass &ps = cp.pass[p];
        const float bias = bf16_bits_to_f32(bias_next);
        bias_next = load_bias(p + 1);
        const uint32_t par = p & 1;
        ptx::mbar_wait(acc_full + 8 * par, (p >> 1) & 1);
        ptx::tc_fence_after_sync();
        if (threadIdx.x == ep_tid0) ft_stamp_pass(cp.trace, p, 3);
        uint32_t own[32];
        ptx::tmem_ld_32x32(tmem_acc + (static_cast<uint32_t>(q * 32) << 16) + par * F2_N + 32 * z, own);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(acc_free + 8 * par);
        const uint32_t recv = smem_recv + par * (S - 1) * F2_RECV_BYTES + (uint32_t)f * 128u;   // this feature's 32 f32
        if (threadIdx.x == ep_tid0) ptx::mbar_arrive_expect_tx(xchg_full + 8 * par, (S - 1) * F2_RECV_BYTES);
        ptx::mbar_wait(xchg_full + 8 * par, (p >> 1) & 1);
        if (threadIdx.x == ep_tid0) ft_stamp_pass(cp.trace, p, 0);
        float v[32];
        if (active) {
          // partial of k-slice z ^ r sits in slot r - 1; summation order own + (z^1) + (z^2) + (z^3): fixed per CTA,
          // hence deterministic (S = 2: IEEE addition is commutative, both CTAs of a pair even round identically)
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(own[j]);
#pragma unroll
          for (int r = 1; r < S; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float4 t;
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                           : "r"(recv + (uint32_t)((r - 1) * F2_RECV_BYTES) + (uint32_t)((j ^ (f & 7)) << 4)));
              v[4 * j] += t.x; v[4 * j + 1] += t.y; v[4 * j + 2] += t.z; v[4 * j + 3] += t.w;
            }
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(recv_free + 8 * par);   // this warp has consumed recv[par] of pass p
        if (active) {
          uint16_t *out = static_cast<uint16_t *>(ps.C) + (int64_t)(m0 + 32 * (int32_t)z) * ps.ldc + n0 + f;
          if (ps.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) out[(int64_t)j * ps.ldc] = f32_to_bf16_bits(relu_f32(v[j] + bias));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) out[(int64_t)j * ps.ldc] = f32_to_bf16_bits(v[j] + bias);
          }
        }
        if (threadIdx.x == ep_tid0) ft_stamp_pass(cp.trace, p, 4);
        if (ps.arrive) asm volatile("bar.arrive 1, 160;" ::: "memory");   // stores issued; the arriver warp publishes them
      }
    }
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  ptx::cluster_sync();                               // the peer may still push to / signal this CTA until it is done too
  if (threadIdx.x == 0) {
    ft_stamp(cp.trace, 2);
    // launch epoch: the last CTA to leave publishes how many barriers every counter has completed
    const unsigned int n_ctas = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int ticket = atomicAdd(cp.epoch + FT_MAX_WAYS, 1u);
    if (ticket == n_ctas - 1) {
#pragma unroll
      for (int sl = 0; sl < FT_MAX_WAYS; ++sl) cp.epoch[sl] += cp.arrivals_total[sl];
      cp.epoch[FT_MAX_WAYS] = 0;
      __threadfence();
    }
  }
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(tmem_acc, 2 * F2_N);
  }
}


} // namespace

// Feature-major chain (mlp_chain_ft_kernel): additionally needs m % 32 == 0, n % 64 == 0, a reduction of exactly
// FT_KB k-blocks per layer, bias-add (bcast_col) or no binary, and (m/32) x (n/64) <= 148 co-resident CTAs.
static bool chain_ft_supported(const KernelDesc *const *descs, const GemmArgs *args, int L) {
  static const bool off = [] { const char *e = getenv("TPP_XSMM_CHAIN"); return e && e[0] == 's'; }();
  if (off) return false;
  const KernelDesc &d0 = *descs[0];
  if ((d0.m % FT_N) != 0 || (d0.n % FT_M) != 0) return false;
  if ((d0.m / FT_N) * (d0.n / FT_M) > 148) return false;
  for (int l = 0; l < L; ++l) {
    const KernelDesc &d = *descs[l];
    const int64_t k_iters = d.k / BLOCK_K, iters = args[l].batch * k_iters;
    if (iters != FT_KB) return false;
    // a group of 4 k-block slots must be one TMA box: 4 k-blocks of one batch element, or whole batch elements
    if (!((k_iters % FT_GROUP) == 0 || k_iters == 1 || k_iters == 2)) return false;
    if (d.op == OpClass::FusedBrgemm && d.binary_kind != 0 && args[l].D == nullptr) return false;
  }
  return true;
}

// The GEN instantiations of the full-K kernel: layers that are grids of tile BRGEMMs on block-packed operands and / or
// have VNNI-2 weights (the reference's default --tiles=32,32,32 --vnni=2 stream after the runtime has folded the tile
// invokes into layers). Same tiling as the flat kernel - 32 batch rows x 64 features per CTA, a reduction of exactly
// FT_KB k-blocks - so every layer is 1024 wide; tile sizes must let one TMA box be a whole number of blocks.
static bool chain_ftg_supported(const KernelDesc *const *descs, const GemmArgs *args, int L, bool *vnni_out, bool *narrow_out) {
  // TPP_XSMM_CHAIN_FTG=0 (read per capture, so that tests can switch it): such chains stay on the pair-per-chain kernel
  const char *env = getenv("TPP_XSMM_CHAIN_FTG");
  static const bool chains_off = [] { const char *e = getenv("TPP_XSMM_CHAIN"); return e && (e[0] == 's' || e[0] == '0'); }();
  if ((env && env[0] == '0') || chains_off || L < 2 || L > CHAIN_MAX_LAYERS) return false;
  {
    // flat chains with flat weights belong to the flat kernels (or to whatever the caller tries next)
    bool special = (descs[0]->gemm_flags & 2048) != 0;
    for (int l = 0; l < L; ++l) special = special || args[l].is_grid();
    if (!special) return false;
  }
  const KernelDesc &d0 = *descs[0];
  const int64_t rows = (int64_t)args[0].grid_n * d0.m, n_total0 = (int64_t)args[0].grid_k * d0.n;
  if ((rows % FT_N) != 0 || (n_total0 % FT_M) != 0 || (rows / FT_N) * (n_total0 / FT_M) > 148) return false;
  const bool vnni = (d0.gemm_flags & 2048) != 0, narrow = d0.k == 32;
  for (int l = 0; l < L; ++l) {
    const KernelDesc &d = *descs[l];
    const GemmArgs &g = args[l];
    if (!brgemm_layer_chainable(d, g)) return false;
    if (((d.gemm_flags & 2048) != 0) != vnni || (d.k == 32) != narrow) return false;   // one instantiation per launch
    if ((int64_t)g.grid_n * d.m != rows || (int64_t)g.grid_k * d.n != n_total0) return false;
    if (g.batch * d.k != (int64_t)FT_KB * BLOCK_K) return false;
    if ((d.m % FT_N) != 0) return false;                                                 // a CTA's 32 rows lie in one row block
    if (!(d.k == 32 || d.k == 64 || d.k == 128 || (d.k % 256) == 0)) return false;       // a group of 4 k-blocks is one box
    if (vnni ? !(d.n == 32 || (d.n % FT_M) == 0) : (d.n % FT_M) != 0) return false;
    if ((d.lda % 8) != 0 || (d.ldb % 8) != 0 || (d.ldc % 8) != 0) return false;
    if (g.batch > 1 && ((d.stride_a % 8) != 0 || (d.stride_b % 8) != 0)) return false;
    if (g.grid_n > 1 && ((g.a_step % 8) != 0 || (g.c_step_n % 8) != 0)) return false;
    if (g.grid_k > 1 && ((g.b_step % 8) != 0 || (g.c_step_k % 8) != 0)) return false;
    if (d.op == OpClass::FusedBrgemm && d.binary_kind != 0 && g.D == nullptr) return false;
  }
  *vnni_out = vnni;
  *narrow_out = narrow;
  return true;
}

// ---- VNNI-2 weights of a lone chain: one flat copy per graph launch ---------------------------------------------------
// The in-kernel converter warps double the shared-memory traffic of a pass (raw rows in, rewritten rows out), and in the
// pass kernels shared memory is the bound: measured +4.4 us per forward (18.4 against 14.0 us), every CTA of a feature
// tile's eight row tiles redoing the same rewrite. Weights do not change inside a chain (hazard-tested), so the launcher
// instead puts ONE small kernel in front of the chain kernel that un-interleaves (and un-blocks) every layer's weights
// into a graph-owned scratch [K][N] (vnni_flat.cu) - 12 MiB of L2-resident traffic per graph launch, shared by all exact
// repeats of the chain in that launch - and the chain kernel reads the flat copy through the ordinary weight map. Every
// replay of the graph converts again: weights the caller changed between two launches are seen.
// TPP_XSMM_FT_VNNI_SCRATCH=0 keeps the converter warps. VNNI-4 weights (mlir-gen --vnni=4) always take the flat copy: the
// converter warps rewrite factor 2 only.
// tensor maps of a GEN pass (see the kernel's header comment); sizes in elements. w_flat: a flat [K][N] copy of the
// layer's VNNI-2 weights (above) to read instead of g.B
static bool encode_ftg_maps(FtPass &ps, const KernelDesc &d, const GemmArgs &g, bool vnni, const void *w_flat = nullptr) {
  const uint64_t nb = (uint64_t)g.batch, gn = (uint64_t)g.grid_n, gk = (uint64_t)g.grid_k;
  // a dimension of size 1 may carry any legal stride
  const uint64_t sa = nb > 1 ? (uint64_t)d.stride_a : (uint64_t)d.lda, sb = nb > 1 ? (uint64_t)d.stride_b : (uint64_t)d.ldb;
  const uint64_t a_step = gn > 1 ? (uint64_t)g.a_step : (uint64_t)d.lda, b_step = gk > 1 ? (uint64_t)g.b_step : (uint64_t)d.ldb;
  const uint32_t kx = (uint32_t)std::min<int64_t>(d.k, BLOCK_K);          // 32 or 64
  const uint32_t kbs = (uint32_t)std::max<int64_t>(d.k / BLOCK_K, 1);     // k-blocks per batch element
  const uint32_t box_kb = std::min<uint32_t>(kbs, FT_GROUP);              // k-blocks of one batch element per box
  const uint32_t box_be = FT_GROUP * BLOCK_K / (kx * box_kb);             // batch elements per box: 256 k in all
  {
    const uint64_t dims[5] = {kx, (uint64_t)d.m, kbs, nb, gn}, str[4] = {(uint64_t)d.lda, BLOCK_K, sa, a_step};
    const uint32_t box[5] = {kx, FT_N, box_kb, box_be, 1};
    if (!encode_map_nd(&ps.tmX, g.A, 5, dims, str, box, kx == 32 ? 64 : 128)) return false;
  }
  ps.w_n = (int32_t)d.n;
  if (w_flat) {
    // the flat copy: one [K][N] matrix, 64 features x 256 k per box
    const uint64_t n_total = gk * (uint64_t)d.n, k_total = nb * (uint64_t)d.k;
    const uint64_t dims[4] = {n_total, k_total, 1, 1}, str[3] = {n_total, n_total, n_total};
    const uint32_t box[4] = {FT_M, FT_GROUP * BLOCK_K, 1, 1};
    if (!encode_map_nd(&ps.tmW, w_flat, 4, dims, str, box, 128)) return false;
    ps.w_n = (int32_t)n_total;
  } else if (vnni) {
    // raw VNNI-2 rows: (element of the [n][2] row | column block | k pair | batch element); 64 features x 2 = 256 bytes
    // per k pair (two column blocks when n == 32), 128 k pairs per box; no swizzle
    const uint32_t ex = 2 * (uint32_t)std::min<int64_t>(d.n, FT_M), kpx = (uint32_t)std::min<int64_t>(d.k / 2, FT_GROUP * BLOCK_K / 2);
    const uint64_t dims[4] = {2 * (uint64_t)d.n, gk, (uint64_t)d.k / 2, nb}, str[3] = {b_step, 2 * (uint64_t)d.ldb, sb};
    const uint32_t box[4] = {ex, 2 * FT_M / ex, kpx, FT_GROUP * BLOCK_K / 2 / kpx};
    if (!encode_map_nd(&ps.tmW, g.B, 4, dims, str, box, 0)) return false;
  } else {
    const uint32_t kw = (uint32_t)std::min<int64_t>(d.k, FT_GROUP * BLOCK_K);
    const uint64_t dims[4] = {(uint64_t)d.n, (uint64_t)d.k, nb, gk}, str[3] = {(uint64_t)d.ldb, sb, b_step};
    const uint32_t box[4] = {FT_M, kw, FT_GROUP * BLOCK_K / kw, 1};
    if (!encode_map_nd(&ps.tmW, g.B, 4, dims, str, box, 128)) return false;
  }
  ps.m = (int32_t)d.m;
  ps.n = (int32_t)d.n;
  ps.c_step_n = gn > 1 ? g.c_step_n : 0;
  ps.c_step_k = gk > 1 ? g.c_step_k : 0;
  for (int gq = 0; gq < FT_NG; ++gq) {
    const int64_t k0 = (int64_t)gq * FT_GROUP * BLOCK_K;     // first reduction index of the group
    ps.grp_be[gq] = (int16_t)(k0 / d.k);
    ps.grp_kb[gq] = (int16_t)((k0 % d.k) / BLOCK_K);
    ps.w_be[gq] = w_flat ? (int16_t)0 : ps.grp_be[gq];
    ps.w_kb[gq] = w_flat ? (int16_t)(k0 / BLOCK_K) : ps.grp_kb[gq];
  }
  return true;
}

// split-K variants (mlp_chain_fts_kernel<S>): 32 S-row batch tiles, a multiple of S feature tiles, and a reduction
// whose S slices are made of whole TMA boxes. Returns the largest usable S in {4, 2}, or 1.
static int chain_ft_split(const KernelDesc *const *descs, const GemmArgs *args, int L) {
  // S = 4 is implemented and parity-clean but slower than S = 2 (7.97 vs 5.71 us per forward): its 24 KiB of st.async
  // pushes per pass move at ~8 B/clk and become the bound. TPP_XSMM_CHAIN_SPLIT=4 enables it, =1 disables split-K.
  static const int max_split = [] { const char *e = getenv("TPP_XSMM_CHAIN_SPLIT"); return e ? atoi(e) : 2; }();
  const KernelDesc &d0 = *descs[0];
  for (int S = 4; S >= 2; S /= 2) {
    if (S > max_split) continue;
    const int rows = 32 * S, kb = FT_KB / S, group = kb / 4;
    if ((d0.m % rows) != 0 || ((d0.n / FT_M) % S) != 0) continue;
    if ((d0.m / rows) * (d0.n / FT_M) * S > 148) continue;
    bool ok = true;
    for (int l = 0; l < L && ok; ++l) {
      const int64_t k_iters = descs[l]->k / BLOCK_K;
      if (!(k_iters == 1 || (k_iters % group) == 0)) ok = false;            // a box = `group` k-blocks of one batch element
      if (k_iters > kb && (k_iters % kb) != 0) ok = false;                   // a k-slice divides a batch element ...
      if (k_iters < kb && (kb % k_iters) != 0) ok = false;                   // ... or is whole batch elements
    }
    (void)args;
    if (ok) return S;
  }
  return 1;
}


// Launch chains [0, num_chains) - chain c is layers [first[c], first[c] + len[c]) of descs / args, each already accepted
// by brgemm_chain_supported - as ONE feature-major launch, interleaving pairs of chains. Only a prefix of mutually
// independent, identically tiled chains is taken. Returns the number of chains launched (0: not applicable).
int launch_brgemm_chains_ft(const KernelDesc *const *descs, const GemmArgs *args, const int *first, const int *len,
                            int num_chains, cudaStream_t stream) {
  if (num_chains < 1) return 0;
  // flat chains: the kernels of rounds 1-2; grids of tile invokes / VNNI-2 weights: the GEN instantiations (full K only)
  bool gen = false, gen_vnni = false, gen_narrow = false;
  if (!brgemm_chain_supported(descs + first[0], args + first[0], len[0]) ||
      !chain_ft_supported(descs + first[0], args + first[0], len[0])) {
    if (!chain_ftg_supported(descs + first[0], args + first[0], len[0], &gen_vnni, &gen_narrow)) return 0;
    gen = true;
  }
  auto chain_ok = [&](int c) {
    if (!gen)
      return brgemm_chain_supported(descs + first[c], args + first[c], len[c]) &&
             chain_ft_supported(descs + first[c], args + first[c], len[c]);
    bool v = false, nr = false;
    return chain_ftg_supported(descs + first[c], args + first[c], len[c], &v, &nr) && v == gen_vnni && nr == gen_narrow;
  };
  static const bool multi_off = [] { const char *e = getenv("TPP_XSMM_CHAIN_MULTI"); return e && e[0] == '0'; }();
  const KernelDesc &d0 = *descs[first[0]];
  const int64_t rows0 = (int64_t)args[first[0]].grid_n * d0.m, ncols0 = (int64_t)args[first[0]].grid_k * d0.n;
  int split = gen ? 1 : chain_ft_split(descs + first[0], args + first[0], len[0]);
  // ---- which chains go into this launch ----
  // `sequential`: the following chains are EXACT REPEATS of the first one (same descriptors, same buffers: the unrolled
  // iterations of a benchmark loop, lib/TPP/Runner/MLIRBench.cpp:265-300). They depend on each other, but only through
  // buffers every CTA group (one 32-row batch tile) partitions the same way, so they can share one launch as a plain
  // sequence of passes: a repeat's first layer reads the never-written input (no wait, it overlaps the previous
  // repeat's last layer), its stores to an intermediate buffer come after every CTA of the row tile has passed the
  // barrier behind the previous repeat's reader of that buffer (barriers are cumulative and no CTA runs more than one
  // arrival ahead), and the final output is rewritten by the same CTA in order. What this removes is the ~5 us between
  // two graph launches that a one-forward-per-graph replay pays.
  bool sequential = false;
  auto is_repeat = [&](int c) {
    if (len[c] != len[c - 1]) return false;
    for (int l = 0; l < len[c]; ++l) {
      const GemmArgs &a = args[first[c] + l], &b = args[first[c - 1] + l];
      if (descs[first[c] + l] != descs[first[c - 1] + l] || a.A != b.A || a.B != b.B || a.C != b.C || a.D != b.D ||
          a.batch != b.batch)
        return false;
    }
    return true;
  };
  int take = 1, passes = len[0];
  {
    std::vector<ByteRange> in_all, out_all;
    chain_ranges(descs + first[0], args + first[0], len[0], in_all, out_all);
    while (!multi_off && take < num_chains) {
      const int c = take;
      const KernelDesc &d = *descs[first[c]];
      if ((int64_t)args[first[c]].grid_n * d.m != rows0 || (int64_t)args[first[c]].grid_k * d.n != ncols0 ||
          passes + len[c] > FT_MAX_PASSES)
        break;
      if (!chain_ok(c)) break;
      if (!gen && chain_ft_split(descs + first[c], args + first[c], len[c]) != split) break;
      std::vector<ByteRange> in, out;
      chain_ranges(descs + first[c], args + first[c], len[c], in, out);
      bool indep = true;
      for (const ByteRange &o : out) {
        for (const ByteRange &x : in_all) indep = indep && !overlaps(o, x);
        for (const ByteRange &x : out_all) indep = indep && !overlaps(o, x);
      }
      for (const ByteRange &i : in)
        for (const ByteRange &x : out_all) indep = indep && !overlaps(i, x);
      if (!indep) {
        if ((take == 1 || sequential) && is_repeat(c)) sequential = true;   // a run of exact repeats
        else break;
      } else if (sequential) {
        break;                                                            // do not mix the two kinds in one pass list
      }
      in_all.insert(in_all.end(), in.begin(), in.end());
      out_all.insert(out_all.end(), out.begin(), out.end());
      passes += len[c];
      ++take;
    }
  }
  // more than a handful of independent block-packed / VNNI-2 chains: the pair-per-chain kernel with column-split items is
  // faster (measured: ~47 us per launch up to 18 row blocks against ~7 us per chain here); exact repeats stay here
  constexpr int kFtGenMax = 6;
  if (gen && !sequential && take > kFtGenMax) return 0;
  // one or two chains per launch have nothing to hide the exchange latency behind: the full-K kernel is faster there
  // (11.0 vs 14.7 us for a single forward); every split-K shape is also a full-K shape
  if (take < 3 || sequential) split = 1;
  const bool split2 = split > 1;
  // ---- the pass list: `ways` chains at a time interleaved layer by layer; a chain's counter slot is its position in
  // the tuple. One chain's layer-to-layer latency (store, fence, counter, poll, TMA: ~5000 clk) is longer than one
  // pass (~3000-4000 clk), so three chains are needed to keep the tensor pipe busy. ----
  static const int ways = [] {
    const char *e = getenv("TPP_XSMM_CHAIN_WAYS");
    const int w = e ? atoi(e) : 3;
    return w < 1 ? 1 : w > FT_MAX_WAYS ? FT_MAX_WAYS : w;
  }();
  static FtParams cp;   // ~20 KiB: too large for the stack of a small thread; launches are serialised per thread anyway
  static std::mutex cp_mutex;
  std::lock_guard<std::mutex> lock(cp_mutex);
  memset(&cp, 0, sizeof(cp));
  uint32_t arrivals[FT_MAX_WAYS] = {0, 0, 0, 0};
  int np = 0;
  bool weights_early = true;
  // VNNI-2 weights: flat copies made by one kernel in front of this one (see vnni2_weights_to_flat_kernel); a weight
  // buffer shared by several passes (exact repeats) is converted once. More than 16 distinct buffers, or column blocks /
  // row pitches the 16-byte path cannot take: the converter warps of the <VNNI> instantiations do the job instead.
  static const bool scratch_off = [] { const char *e = getenv("TPP_XSMM_FT_VNNI_SCRATCH"); return e && e[0] == '0'; }();
  // (one conversion kernel per graph launch pays off when the launch re-reads the weights - the exact repeats of an unrolled
  // loop: 10.2 against 15.7 us per forward; a single forward per launch is faster with the converter warps, 18.4
  // against 20.5 us: the extra kernel node costs more than the rewrite it saves)
  const bool vnni4 = gen && gen_vnni && descs[first[0]]->vnni_factor == 4;
  bool w_scratch = gen && gen_vnni && (vnni4 || (!scratch_off && sequential));
  std::vector<VnniFlatJob> wf;
  if (w_scratch) {
    for (int c = 0; c < take && w_scratch; ++c)
      for (int l = 0; l < len[c] && w_scratch; ++l) {
        const KernelDesc &d = *descs[first[c] + l];
        const GemmArgs &g = args[first[c] + l];
        bool seen = false;
        for (const VnniFlatJob &e : wf) seen = seen || e.src == g.B;
        if (seen) continue;
        if (!vnni_flat_job_ok(d, g)) { w_scratch = false; break; }
        wf.push_back(vnni_flat_job(d, g, nullptr));
      }
    if (!w_scratch && vnni4) return 0;   // no converter warps for factor 4: somebody else's chain
    if (w_scratch)
      for (VnniFlatJob &wl : wf) {
        void *buf = nullptr;
        TPP_CUDA_CHECK(cudaMalloc(&buf, (size_t)wl.nb * wl.k * wl.gk * wl.n * sizeof(uint16_t)));
        capture_adopt(buf);   // owned by the graph being captured
        wl.dst = buf;
      }
  }
  auto flat_copy_of = [&](const void *B) -> const void * {
    for (const VnniFlatJob &e : wf)
      if (e.src == B) return e.dst;
    return nullptr;
  };
  auto add_pass = [&](int c, int l, int slot) -> bool {
    const KernelDesc &d = *descs[first[c] + l];
    const GemmArgs &g = args[first[c] + l];
    FtPass &ps = cp.pass[np];
    const uint64_t nb = (uint64_t)g.batch;
    const uint32_t k_iters = (uint32_t)(d.k / BLOCK_K);
    const uint32_t grp = split == 4 ? FS<4>::GROUP : split == 2 ? FS<2>::GROUP : FT_GROUP;
    const uint32_t gk = k_iters >= grp ? grp : (k_iters ? k_iters : 1), gb = grp / gk;   // box = gk k-blocks x gb batch elements
    if (gen) {
      if (!encode_ftg_maps(ps, d, g, gen_vnni, w_scratch ? flat_copy_of(g.B) : nullptr)) return false;
    } else if (!encode_map_x4(&ps.tmX, g.A, (uint64_t)d.k, (uint64_t)d.m, nb, (uint64_t)d.lda, (uint64_t)d.stride_a,
                              32 * split, gk, gb) ||
               !encode_map(&ps.tmW, g.B, (uint64_t)d.n, (uint64_t)d.k, nb, (uint64_t)d.ldb, (uint64_t)d.stride_b, FT_M,
                           BLOCK_K * gk, gb))
      return false;
    ps.C = g.C;
    ps.D = g.D;
    ps.ldc = d.ldc;
    ps.k_iters = (int32_t)k_iters;
    ps.groups = FT_NG;
    ps.has_bias = (d.op == OpClass::FusedBrgemm && g.D && d.binary_kind == 1) ? 1 : 0;
    ps.relu = (d.op == OpClass::FusedBrgemm && d.unary_kind == 5) ? 1 : 0;
    ps.slot = (uint8_t)slot;
    ps.x_dep = l > 0 ? 1 : 0;
    ps.wait_arrivals = arrivals[ps.slot];
    ps.arrive = l + 1 < len[c] ? 1 : 0;
    if (ps.arrive) ++arrivals[ps.slot];
    if (!g.b_independent || w_scratch) weights_early = false;   // the flat copies are written by the kernel just before
    ++np;
    return true;
  };
  bool ok = true;
  if (sequential) {
    for (int c = 0; c < take && ok; ++c)
      for (int l = 0; l < len[c] && ok; ++l) ok = add_pass(c, l, 0);
  } else
  for (int c = 0; c < take && ok; c += ways) {
    const int nc = std::min(ways, take - c);
    int maxL = 0;
    for (int j = 0; j < nc; ++j) maxL = std::max(maxL, len[c + j]);
    for (int l = 0; l < maxL && ok; ++l)
      for (int j = 0; j < nc && ok; ++j)
        if (l < len[c + j]) ok = add_pass(c + j, l, j);
  }
  if (!ok) {
    static bool warned = false;
    if (!warned) fprintf(stderr, "tpp-xsmm-cuda: feature-major chain: tensor map encode failed, using the split-K chain\n");
    warned = true;
    return 0;
  }
  dim3 grid((unsigned)(ncols0 / FT_M), (unsigned)(rows0 / (32 * split)), (unsigned)split);
  const int n_ctas = (int)(grid.x * grid.y * grid.z);
  // arrival counters of THIS kernel node (the chain kernels only ever run inside a capture): zero-filled before the
  // node exists, owned by the graph, monotonic across its replays - every counter stays a multiple of the group size
  // between launches; the split-K-2 variant's launch epoch (+ exit ticket) lives behind its counters
  unsigned int *counters = static_cast<unsigned int *>(
      capture_owned_zeroed(sizeof(unsigned int) * (FT_MAX_WAYS * FT_CTR_SLOT + 2 * FT_MAX_WAYS)));
  cp.counters = counters;
  cp.epoch = counters + FT_MAX_WAYS * FT_CTR_SLOT;
  for (int sl = 0; sl < FT_MAX_WAYS; ++sl) cp.arrivals_total[sl] = arrivals[sl];
  cp.num_passes = np;
  cp.weights_early = weights_early ? 1 : 0;
  if (getenv("TPP_XSMM_DEBUG"))
    fprintf(stderr, "ft-chain: %d chains, %d passes, gen=%d vnni=%d narrow=%d sequential=%d weights_early=%d a_independent=%d\n", take, np,
            (int)gen, (int)gen_vnni, (int)gen_narrow, (int)sequential, (int)weights_early, (int)args[first[0]].a_independent);
  static const bool x0_off = [] { const char *e = getenv("TPP_XSMM_CHAIN_X0"); return e && e[0] == '0'; }();
  cp.x0_early = (args[first[0]].a_independent && !x0_off) ? 1 : 0;
  // fence.proxy.async between the flag observation and the TMA reads costs ~0.3 us per layer and is not needed for
  // data that other SMs fenced to L2 (TMA reads L2); TPP_XSMM_CHAIN_PROXY_FENCE=1 turns it on
  static const bool pf = [] { const char *e = getenv("TPP_XSMM_CHAIN_PROXY_FENCE"); return e && e[0] == '1'; }();
  cp.proxy_fence = pf ? 1 : 0;
  constexpr int smem1 = FT_KB * (FT_X_BYTES + FT_W_BYTES) + FT_OUT_BYTES + (4 * FT_NG + 4) * 8 + 16 + 1024;
  const int smem = split == 4 ? FS<4>::SMEM : split == 2 ? FS<2>::SMEM : smem1;
  // the kernel this launch runs (full-K instantiations; the split-K variants are set below)
  using FtKernel = void (*)(const FtParams);
  const bool conv_warps = gen && gen_vnni && !w_scratch;   // VNNI-2 weights rewritten inside the kernel
  const FtKernel ft_kernel = !gen                        ? mlp_chain_ft_kernel<false, false, false>
                             : conv_warps && gen_narrow ? mlp_chain_ft_kernel<true, true, true>
                             : conv_warps               ? mlp_chain_ft_kernel<true, true, false>
                             : gen_narrow             ? mlp_chain_ft_kernel<true, false, true>
                                                      : mlp_chain_ft_kernel<true, false, false>;
  static std::once_flag once;
  std::call_once(once, [] {
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_ft_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_ft_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_ft_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_ft_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_ft_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1));
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_fts_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS<2>::SMEM));
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_fts_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, FS<4>::SMEM));
  });
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(split2 ? F2_THREADS : conv_warps ? FT_THREADS_VNNI : NUM_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[3];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  // TPP_XSMM_CHAIN_MC=1: weight multicast across pairs of batch tiles ((1,2,1) clusters). It halves the L2 reads of
  // the weights but not the bytes each SM receives, and a pass is bound by the latter (~47-53 B/clk per SM):
  // measured 7.15 us (multicast) vs 7.12 us (unicast) per forward, so it stays off by default.
  static const bool mc_on = [] { const char *e = getenv("TPP_XSMM_CHAIN_MC"); return e && e[0] == '1'; }();
  cp.w_multicast = (!split2 && !gen && mc_on && (grid.y % 2) == 0) ? 1 : 0;
  if (cp.w_multicast || split2) {
    attrs[1].id = cudaLaunchAttributeClusterDimension;
    attrs[1].val.clusterDim.x = 1;
    attrs[1].val.clusterDim.y = split2 ? 1 : 2;
    attrs[1].val.clusterDim.z = split2 ? (unsigned)split : 1;
    cfg.numAttrs = 2;
  }
  static const bool trace_on = [] { const char *e = getenv("TPP_XSMM_TC_TRACE"); return e && atoi(e) == 3; }();
  if (trace_on) {
    if (!g_trace_buf) {
      TPP_CUDA_CHECK(cudaMalloc(&g_trace_buf, sizeof(unsigned long long) * kTraceRing * kTraceRingCtas * TRACE_SLOTS));
      TPP_CUDA_CHECK(cudaMemsetAsync(g_trace_buf, 0, sizeof(unsigned long long) * kTraceRing * kTraceRingCtas * TRACE_SLOTS, stream));
    }
    cp.trace = g_trace_buf;
    g_chain_trace_ctas = n_ctas;
    g_chain_trace_layers = np;
    g_chain_trace_ft = true;
  }
  // the passes wait for each other through arrival counters: every CTA must be resident (cooperative launch)
  const void *kfn = split == 4 ? reinterpret_cast<const void *>(mlp_chain_fts_kernel<4>)
                    : split == 2 ? reinterpret_cast<const void *>(mlp_chain_fts_kernel<2>)
                                 : reinterpret_cast<const void *>(ft_kernel);
  if (!prepare_resident_launch(kfn, &cfg, attrs)) return 0;
  if (w_scratch) launch_vnni_weights_to_flat(wf.data(), (int)wf.size(), stream);
  if (split == 4) TPP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mlp_chain_fts_kernel<4>, cp));
  else if (split == 2) TPP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mlp_chain_fts_kernel<2>, cp));
  else TPP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, ft_kernel, cp));
  const char *tile = split == 4 ? "ft64x128_splitk4" : split == 2 ? "ft64x64_splitk2" : "ft64x32_fullk";
  char tag[24] = "";
  if (gen) snprintf(tag, sizeof(tag), "%s%s", args[first[0]].is_grid() ? "_blocked" : "", vnni4 ? "_vnni4" : gen_vnni ? "_vnni2" : "");
  if (take == 1) set_last_name("mlp_chain_bf16_%dlayers_%s%s", len[0], tile, tag);
  else set_last_name("mlp_chain_bf16_%dx%dlayers_%s%s%s", take, len[0], tile, tag, sequential ? "_seq" : "");
  return take;
}


} // namespace tpp
