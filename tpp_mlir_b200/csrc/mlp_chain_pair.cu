// mlp_chain_pair.cu - pair-per-chain kernel: one CTA pair (tcgen05.mma.cta_group::2, 256 x 256 tiles) walks a whole
// layer chain for a block of 256 batch rows, many chains side by side. The dominant kernel of bench.py. DESIGN.md 4.1d.
#include "tc_common.cuh"

namespace tpp {
using namespace tc;

namespace {

// ---- pair-per-chain kernel: one CTA pair runs a whole layer chain, many chains side by side ---------------------------
// Third design of the fused chain, for launches that carry MANY independent chains (a captured graph of the
// benchmark's rotating operand sets, a batch of requests, or the 256-row blocks of a large-batch MLP: rows are
// independent through all layers). The pass kernels above spread ONE layer over 128 SMs in 64 x 64 tiles: every SM then
// receives 128 KiB of operands per layer pass, 16 MiB per layer over all SMs for 2.5 MiB of unique data, and the
// chip-wide L2 -> SM throughput (~6300 B/clk) bounds a pass at ~2700 clk no matter how the latencies are hidden.
// Here a work item (one chain x one block of 256 batch rows) belongs to ONE pair of CTAs (two SMs of a TPC,
// tcgen05.mma.cta_group::2, M = 256) which walks the layers and, per layer, the 256-column output tiles:
//   * per tile and k-block each CTA stages its 128 activation rows (16 KiB) and HALF of the 256 weight columns (16 KiB);
//     a layer costs 4 MiB of L2 -> SM traffic per item instead of 16 MiB, every weight byte is fetched exactly once;
//   * CTA r only ever reads the activation rows it wrote itself (rows 128 r .. 128 r + 127 of the item), so a layer
//     boundary needs no cross-SM synchronisation at all: the epilogue thread that issues the CTA's TMA stores waits
//     for their completion and arrives on a LOCAL mbarrier per output tile; the CTA's producer waits for tile i / 4
//     before reduction step i of the next layer's first tile. No counters, no co-residency assumption, nothing to
//     spin on across SMs; the last epilogue of a layer hides behind 12 of the next tile's 16 reduction steps;
//   * the next layer's weights do not depend on anything: their box of a ring slot is always issued BEFORE that wait
//     (same mbarrier, expect_tx covers both operands);
//   * TMEM holds two 256-column accumulators: the epilogue of tile t (tcgen05.ld -> bias from shared memory -> ReLU ->
//     bf16 -> swizzled staging buffer -> TMA store) runs under the MMAs of tile t + 1;
//   * L2 eviction-priority hints keep the activations (re-read once per output tile) resident under the weight stream;
//   * pairs are independent: the grid is min(items, 74) pairs, pair p takes items p, p + pairs, ...
// Layer descriptors (three tensor maps + epilogue parameters per layer) live in a device table written once at capture.
// History of the measurements that shaped it: profiles/kernel_trace_r1.txt.
constexpr int PC_STAGES = 6;
constexpr int PC_BLOCK_N = 256;                       // output columns per tile (UMMA N)
constexpr int PC_HALF_N = PC_BLOCK_N / 2;             // weight columns staged by each CTA
constexpr int PC_W_CHUNKS = PC_HALF_N / 64;           // 64-column TMA boxes per CTA and k-block
constexpr int PC_STAGE_BYTES = A_STAGE_BYTES + PC_W_CHUNKS * B_CHUNK_BYTES;   // 32 KiB
constexpr int PC_ROWS = 2 * BLOCK_M;                  // batch rows per work item
constexpr int PC_OUT_COLS = 64;                       // columns per TMA store box (128 bytes: one swizzle row)
constexpr int PC_OUT_BYTES = BLOCK_M * PC_OUT_COLS * 2;   // 16 KiB staging buffer, two of them
constexpr int PC_BIAS_BYTES = PC_BLOCK_N * 2;         // one tile's bias slice, two of them
constexpr int PC_MAX_TILES = 16;                      // output tiles per layer (n <= 4096): one "stored" barrier each
constexpr int PC_SMEM = PC_STAGES * PC_STAGE_BYTES + 2 * PC_OUT_BYTES + 2 * PC_BIAS_BYTES +
                        (3 * PC_STAGES + 4 + PC_MAX_TILES) * 8 + 16 + 1024;

// Operand addressing. A layer is a grid of tile BRGEMMs on block-packed operands (GemmArgs::grid_*; 1 x 1 with
// m = 256 r, k = K is the flat case): all three operands are described by 4-D tensor maps whose box gathers one
// 128-byte swizzle row from several blocks -
//   X (k in block | batch element | row in block | row block): box (min(k,64), 64/min(k,64), min(m,128), 128/min(m,128))
//   W (n in block | column block | k in block | batch element): box (min(n,64), 64/min(n,64), min(k,64), 64/min(k,64))
//   C (n in block | column block | row in block | row block):   box (min(n,64), 64/min(n,64), min(m,128), 128/min(m,128))
// so shared memory always receives the canonical K-major (X) / MN-major (W) SWIZZLE_128B tiles the MMA descriptors
// expect, whatever the tiling of the caller. 32-wide blocks (the reference's default --tiles=32,32,32) have 64-byte
// innermost extents, which TMA would pad to one 128-byte line each under SWIZZLE_128B (scripts/probes/tma_box_probe.cu):
// those operands use SWIZZLE_64B sub-tiles instead (NARROW instantiations). VNNI-2 weights ([k/2][n][2], the reference's
// default bf16 layout) are not a canonical UMMA layout and TMA cannot de-interleave 2-byte elements: in the
// <VNNI = true> instantiations TMA drops the raw rows into the ring slot and eight converter warps per CTA rewrite them
// IN PLACE into the swizzled MN-major tile (LDS.128 -> 4 PRMT + 2 SHFL -> STS.128) - no extra pass through HBM, no
// extra shared memory; they arrive on the same "full" barrier the activations' TMA bytes are counted on.
struct alignas(128) PcLayer {
  CUtensorMap tmX;          // activations, box = 64 k x 128 rows
  CUtensorMap tmW;          // flat weights, box = 64 n x 64 k (unused for VNNI-2 weights)
  CUtensorMap tmC;          // output, box = 64 n x 128 rows (TMA store from the swizzled staging buffer)
  const void *D;            // bias vector or nullptr
  const void *W;            // VNNI-2 weights: base pointer for the converter warps
  int64_t w_col_step;       // VNNI-2: elements between column blocks (b_step)
  int64_t w_batch_step;     // VNNI-2: elements between batch elements (stride_b)
  int64_t w_ldb;            // VNNI-2: leading dimension in k pairs (elements / 2 per k-pair row = ldb)
  int32_t m, n, k;          // the tile BRGEMM: rows per row block, columns per column block, k per batch element
  int32_t k_bstep;          // batch elements per 64-wide k-block (k < 64), else 1
  int32_t total_iters;      // k-blocks of the whole reduction (batch x k / 64)
  int32_t n_tiles;          // output columns / 256
  int32_t relu;
  int32_t vnni;
  // 32-wide blocks: the innermost contiguous extent of an operand is 64 bytes, half a SWIZZLE_128B row (TMA would pad
  // every 64-byte piece to its own 128-byte line). Those operands use SWIZZLE_64B instead: a 64-wide k-block of X is
  // two [128 rows][32 k] sub-tiles (two boxes), a 64-column chunk of W / C two [64 k | 128 rows][32 n] sub-tiles.
  int32_t x64;              // k == 32
  int32_t w64;              // n == 32, flat weights
  int32_t c64;              // n == 32
  int32_t w_flat;           // tmW describes a flat [K][N] copy of the weights (vnni_flat.cu): coordinates (column, 0, k, 0)
  // The layer's OUTPUT is a function-local temporary (xsmm_cuda_mark_temporary) whose only reader is the next layer of
  // this chain: once a CTA has finished that layer, the 128 output rows it wrote and re-read are dead and it drops their
  // cache lines from L2 (discard.global.L2) - they never travel to HBM. A CTA's rows of the output are d_nrb x d_gk
  // contiguous chunks of d_lines 128-byte lines (block-packed outputs: one chunk per (row block, column block); flat:
  // one chunk); 0 lines: not a temporary / not expressible, nothing is discarded.
  const char *d_base;       // output base address (bytes)
  int64_t d_step_n, d_step_k, d_row_bytes;   // bytes between row blocks / column blocks / rows inside a block
  int32_t d_lines, d_nrb, d_gk;
};
// A work item: rows [row0, row0 + 256) of one chain. In a launch with FEW row blocks (a lone forward pass) the output
// tiles of every layer are dealt out to `nslices` pairs instead - pair `slice` takes tiles slice, slice + nslices, ... - so
// that one chain occupies several SM pairs; the layer boundary then crosses SMs and goes through per-tile flags in
// global memory (flag_base, see PcParams::flags) instead of the CTA-local barriers.
struct PcItem {
  int32_t layer0, num_layers, row0;
  int32_t slice, nslices;   // 0, 1: the whole row block belongs to one pair
  int32_t flag_base;        // first flag word of this row block: [layer][tile][CTA of the pair]
  int32_t pad[2];
};
struct PcParams {
  const PcLayer *layers;
  const PcItem *items;
  int32_t num_items;
  int32_t l2_hints;            // L2 eviction-priority hints on the TMA loads / stores (TPP_XSMM_CHAIN_PAIR_HINTS=0: off)
  // column-split items only: flags[flag_base + (l * PC_MAX_TILES + j) * 2 + r] = epoch + 1 once rows 128 r .. of output
  // tile j of layer l are in L2; epoch[0] = launches completed so far on this table (read at kernel start, bumped by the
  // last CTA to leave - flags are never reset), epoch[1] = exit ticket
  unsigned int *flags;
  unsigned int *epoch;
  int32_t debug;               // TPP_XSMM_PAIR_DEBUG (timing experiments only, results are wrong): 1 = converters skip the
                               // rewrite, 2 = converters skip the proxy fence
  unsigned long long *trace;   // TPP_XSMM_TC_TRACE=4: clock stamps of each CTA's first item (nullptr in normal runs)
};
constexpr int PC_CONV_WARPS = 8;                      // converter warps (VNNI): 256 threads x 4 pieces of 16 bytes = 16 KiB per k-block
constexpr int PC_THREADS_VNNI = NUM_THREADS + 32 * PC_CONV_WARPS;   // 448
__device__ __forceinline__ void pc_stamp(const PcParams &cp, int slot) {
  if (cp.trace) cp.trace[(size_t)blockIdx.x * PC_TRACE_SLOTS + slot] = clock64();
}

__device__ __forceinline__ void tensormap_acquire(const void *map) {
  // the table was written by a host copy: make it visible to the tensor-map proxy of this SM before the first use
  asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// NARROW: some layer has 32-wide blocks (SWIZZLE_64B operands); false compiles the selects out of the hot loops
template <bool VNNI, bool NARROW>
__global__ void __launch_bounds__(VNNI ? PC_THREADS_VNNI : NUM_THREADS, 1) mlp_chain_pair_kernel(const PcParams cp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;                                       // PC_STAGES x 16 KiB
  const uint32_t smem_w = smem_base + PC_STAGES * A_STAGE_BYTES;           // PC_STAGES x 2 x 8 KiB
  const uint32_t smem_out = smem_w + PC_STAGES * PC_W_CHUNKS * B_CHUNK_BYTES;   // 2 x 16 KiB output staging
  const uint32_t smem_bias = smem_out + 2 * PC_OUT_BYTES;                  // 2 x 512 B: bias slice of tile t / t + 1
  const uint32_t bar_base = smem_bias + 2 * PC_BIAS_BYTES;
  const uint32_t full_bar = bar_base;                                      // leader's: both CTAs' bytes land on it
  const uint32_t empty_bar = bar_base + PC_STAGES * 8;                     // per CTA, released by the pair's MMA commits
  const uint32_t acc_full = bar_base + 2 * PC_STAGES * 8;                  // [2] per CTA: accumulator complete
  const uint32_t acc_free = acc_full + 16;                                 // [2] leader's: both epilogues have read it out
  const uint32_t tile_done = acc_free + 16;                                // [PC_MAX_TILES] per CTA: my rows of output tile j are stored
  const uint32_t raw_full = tile_done + 8 * PC_MAX_TILES;                  // [PC_STAGES] per CTA (VNNI): my raw weight bytes have landed
  const uint32_t tmem_slot = raw_full + 8 * PC_STAGES;
  uint8_t *smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t peer = ptx::cluster_ctarank();       // 0 = leader
  const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < PC_STAGES; ++s) {
      // VNNI: besides the producer's expect_tx arrival, the converter warps of BOTH CTAs arrive once their part of the
      // weight tile is in place
      ptx::mbar_init(full_bar + 8 * s, VNNI ? 1 + 2 * PC_CONV_WARPS : 1);
      ptx::mbar_init(empty_bar + 8 * s, 1);
      ptx::mbar_init(raw_full + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(acc_full + 8 * b, 1);
      ptx::mbar_init(acc_free + 8 * b, 8);            // one arrival per epilogue warp of both CTAs
    }
    for (int j = 0; j < PC_MAX_TILES; ++j) ptx::mbar_init(tile_done + 8 * j, 1);   // the thread that issues the TMA stores
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_pair(tmem_slot, 2 * PC_BLOCK_N);  // all 512 columns: two accumulators
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before_sync();
  __syncwarp();
  ptx::cluster_arrive();   // both CTAs' barriers and TMEM exist before any remote signal / pair MMA
  ptx::cluster_wait();
  ptx::tc_fence_after_sync();
  const uint32_t tmem_acc = *tmem_slot_ptr;

  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (threadIdx.x == 0) pc_stamp(cp, 60);

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own rows / own weight columns into own smem, bytes counted on the LEADER =====
    if (lane == 0) {
      const uint32_t leader_full = ptx::mapa(full_bar, 0);
      // L2 residency (measured with ncu before the hints: 1.34 GB of DRAM reads per launch for 1.01 GB of operands -
      // the weight stream evicted activations between their four re-reads): weights are used once -> evict_first;
      // activations are re-read once per output tile -> evict_last until the layer's last tile, whose read demotes them
      const uint64_t pol_first = ptx::l2_policy_evict_first(), pol_last = ptx::l2_policy_evict_last();
      const bool hints = cp.l2_hints != 0;
      int s = 0;
      uint32_t ph = 0, done_ph = 0;                   // done_ph bit j: parity of tile_done[j]'s next phase
      unsigned int epoch1 = 0;                        // the flag value of THIS launch (column-split items)
      if (cp.epoch) {
        asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(epoch1) : "l"(cp.epoch) : "memory");
        ++epoch1;
      }
      for (int item = pair; item < cp.num_items; item += num_pairs) {
        const PcItem it = cp.items[item];
        const int32_t row0 = it.row0 + (int32_t)peer * BLOCK_M;
        for (int l = 0; l < it.num_layers; ++l) {
          const PcLayer *L = cp.layers + it.layer0 + l;
          tensormap_acquire(&L->tmX);
          tensormap_acquire(&L->tmW);
          const int32_t total = L->total_iters, n_tiles = L->n_tiles;
          const int32_t lk = L->k, k_bstep = L->k_bstep;
          const bool w_flat = !VNNI && L->w_flat != 0;   // weights come from a flat [K][N] copy: one "block" as wide as the layer
          const int32_t ln = w_flat ? n_tiles * PC_BLOCK_N : L->n;
          const bool x64 = NARROW && L->x64 != 0, w64 = NARROW && L->w64 != 0;
          // my 128 rows: inside one row block (m >= 128) or 128 / m whole row blocks
          const int32_t xr = L->m >= BLOCK_M ? row0 % L->m : 0, xi = row0 / L->m;
          int32_t ready = 0;                          // output tiles of layer l - 1 (my rows) known to be stored
          const bool split = it.nslices > 1;
          const unsigned int *prev_flags = split && l > 0 ? cp.flags + it.flag_base + (l - 1) * PC_MAX_TILES * 2 + (int)peer : nullptr;
          for (int32_t j = it.slice; j < n_tiles; j += it.nslices) {
            const int32_t wcol = j * PC_BLOCK_N + (int32_t)peer * PC_HALF_N;
            int32_t wn[PC_W_CHUNKS], wj[PC_W_CHUNKS];   // my weight columns: (column in block, column block) per 64-column box
#pragma unroll
            for (int c = 0; c < PC_W_CHUNKS; ++c) {
              const int32_t col = wcol + c * 64;
              wn[c] = ln >= 64 ? col % ln : 0;
              wj[c] = col / ln;
            }
            // (column-split items: other pairs re-read these rows too, nobody knows who is last - keep them resident)
            const uint64_t pol_x = (split || j + 1 < n_tiles) ? pol_last : pol_first;
            int32_t c0 = 0, c1 = 0;                   // k within the batch element, batch element of this k-block
            for (int32_t i = 0; i < total; ++i) {
              ptx::mbar_wait(empty_bar + 8 * s, ph ^ 1);
              // both CTAs' bytes: activations, plus the weights unless the converter warps deliver them
              if (peer == 0) ptx::mbar_arrive_expect_tx(full_bar + 8 * s, VNNI ? 2 * A_STAGE_BYTES : 2 * PC_STAGE_BYTES);
              if (VNNI) {
                // VNNI-2 weights: the raw [k/2][n][2] rows of my 128 columns go into the slot as they are (no swizzle),
                // counted on MY raw barrier; my converter warps rewrite them in place
                ptx::mbar_arrive_expect_tx(raw_full + 8 * s, PC_W_CHUNKS * B_CHUNK_BYTES);
#pragma unroll
                for (int c = 0; c < PC_W_CHUNKS; ++c) {
                  const uint32_t dst = smem_w + (s * PC_W_CHUNKS + c) * B_CHUNK_BYTES;
                  // (element pair in block, column block, k pair in batch element, batch element)
                  if (hints) ptx::tma_load_4d_hint(dst, &L->tmW, raw_full + 8 * s, 2 * wn[c], wj[c], c0 >> 1, c1, pol_first);
                  else ptx::tma_load_4d(dst, &L->tmW, raw_full + 8 * s, 2 * wn[c], wj[c], c0 >> 1, c1);
                }
              } else {
                // the weights depend on nothing: their boxes go out before any wait for the previous layer
#pragma unroll
                for (int c = 0; c < PC_W_CHUNKS; ++c) {
                  const uint32_t dst = smem_w + (s * PC_W_CHUNKS + c) * B_CHUNK_BYTES;
                  const int32_t wk = w_flat ? i * BLOCK_K : c0, wb = w_flat ? 0 : c1;   // k inside the batch element, batch element
                  if (hints) ptx::tma_load_4d_pair_hint(dst, &L->tmW, leader_full + 8 * s, wn[c], wj[c], wk, wb, pol_first);
                  else ptx::tma_load_4d_pair(dst, &L->tmW, leader_full + 8 * s, wn[c], wj[c], wk, wb);
                  if (w64) {   // the chunk's second 32-column block: its own [64 k][32 n] sub-tile
                    if (hints) ptx::tma_load_4d_pair_hint(dst + B_CHUNK_BYTES / 2, &L->tmW, leader_full + 8 * s, 0, wj[c] + 1, c0, c1, pol_first);
                    else ptx::tma_load_4d_pair(dst + B_CHUNK_BYTES / 2, &L->tmW, leader_full + 8 * s, 0, wj[c] + 1, c0, c1);
                  }
                }
              }
              if (l > 0 && j == it.slice) {
                // reduction step i reads columns [64 i, 64 i + 64) of the previous layer's output = its tile i / 4:
                // only the last four steps of the first tile have to wait for the previous layer's last epilogue
                const int32_t need = (i * BLOCK_K) / PC_BLOCK_N;
                while (ready <= need) {
                  const bool last = ready + 1 == (total * BLOCK_K) / PC_BLOCK_N;
                  if (last && item == pair) pc_stamp(cp, 48 + 2 * l);
                  if (!split) {
                    ptx::mbar_wait(tile_done + 8 * ready, (done_ph >> ready) & 1u);
                    done_ph ^= 1u << ready;
                  } else {
                    // tile `ready` of the previous layer was produced by another pair: its flag in global memory
                    unsigned int seen, spins = 0;
                    do {
                      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(prev_flags + ready * 2) : "memory");
                      if (++spins > (1u << 24)) __trap();   // cooperative launch: all pairs are resident; never hang silently
                    } while (seen != epoch1);
                  }
                  if (last && item == pair) pc_stamp(cp, 49 + 2 * l);
                  ++ready;
                  asm volatile("fence.proxy.async;" ::: "memory");
                }
              }
              if (hints)
                ptx::tma_load_4d_pair_hint(smem_a + s * A_STAGE_BYTES, &L->tmX, leader_full + 8 * s, c0, c1, xr, xi, pol_x);
              else
                ptx::tma_load_4d_pair(smem_a + s * A_STAGE_BYTES, &L->tmX, leader_full + 8 * s, c0, c1, xr, xi);
              if (x64) {     // k == 32: the k-block's second batch element is its own [128 rows][32 k] sub-tile
                if (hints)
                  ptx::tma_load_4d_pair_hint(smem_a + s * A_STAGE_BYTES + A_STAGE_BYTES / 2, &L->tmX, leader_full + 8 * s, 0, c1 + 1,
                                             xr, xi, pol_x);
                else
                  ptx::tma_load_4d_pair(smem_a + s * A_STAGE_BYTES + A_STAGE_BYTES / 2, &L->tmX, leader_full + 8 * s, 0, c1 + 1, xr, xi);
              }
              c0 += BLOCK_K;
              if (c0 >= lk) { c0 = 0; c1 += k_bstep; }
              if (++s == PC_STAGES) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader only =====
    if (lane == 0 && peer == 0) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(PC_ROWS, PC_BLOCK_N, /*A K-major*/ 0, /*B MN-major*/ 1);
      const uint16_t pair_mask = 3;
      int s = 0;
      uint32_t ph = 0, t = 0;
      for (int item = pair; item < cp.num_items; item += num_pairs) {
        const PcItem it = cp.items[item];
        for (int l = 0; l < it.num_layers; ++l) {
          const PcLayer *L = cp.layers + it.layer0 + l;
          const int32_t total = L->total_iters, n_tiles = L->n_tiles;
          const bool x64 = NARROW && L->x64 != 0, w64 = NARROW && L->w64 != 0;
          for (int32_t j = it.slice; j < n_tiles; j += it.nslices, ++t) {
            const uint32_t buf = t & 1;
            if (t >= 2) {                              // both epilogues have read tile t - 2 out of this accumulator
              ptx::mbar_wait_cluster(acc_free + 8 * buf, ((t >> 1) - 1) & 1);
              ptx::tc_fence_after_sync();
            }
            const uint32_t acc = tmem_acc + buf * PC_BLOCK_N;
            if (t < 12) pc_stamp(cp, 4 * t);
            for (int32_t i = 0; i < total; ++i) {
              ptx::mbar_wait(full_bar + 8 * s, ph);
              ptx::tc_fence_after_sync();
              const uint32_t a_addr = smem_a + s * A_STAGE_BYTES;
              const uint32_t b_addr = smem_w + s * PC_W_CHUNKS * B_CHUNK_BYTES;
#pragma unroll
              for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
                // SWIZZLE_128B: A rows are 128 bytes (64 k), B atoms 64 columns wide. SWIZZLE_64B (32-wide blocks): A is two
                // [128 rows][32 k] sub-tiles (k steps 0,1 | 2,3), 8-row atoms of 512 bytes; B atoms are 32 columns wide
                // (4 KiB apart), 8 k rows = 512 bytes, one k step = 16 rows of 64 bytes
                const uint64_t da = x64 ? ptx::umma_smem_desc_sw64(a_addr + (kk >> 1) * (A_STAGE_BYTES / 2) + (kk & 1) * (UMMA_K * 2), 16, 512)
                                        : ptx::umma_smem_desc_sw128(a_addr + kk * (UMMA_K * 2), 16, 1024);
                const uint64_t db = w64 ? ptx::umma_smem_desc_sw64(b_addr + kk * (UMMA_K * 64), B_CHUNK_BYTES / 2, 512)
                                        : ptx::umma_smem_desc_sw128(b_addr + kk * (UMMA_K * 128), B_CHUNK_BYTES, 1024);
                ptx::umma_bf16_pair(acc, da, db, idesc, (i > 0 || kk > 0) ? 1u : 0u);
              }
              ptx::umma_commit_pair(empty_bar + 8 * s, pair_mask);   // frees the slot in both CTAs
              if (++s == PC_STAGES) { s = 0; ph ^= 1; }
            }
            ptx::umma_commit_pair(acc_full + 8 * buf, pair_mask);    // both epilogues may start
            if (t < 12) pc_stamp(cp, 4 * t + 1);
          }
        }
      }
    }
  } else if (warp < 6) {
    // ===== epilogue (both CTAs, each on its own 128 rows / TMEM lanes) =====
    // TMEM lane = row: a thread owns one output row. Its bf16 results go to a 128-byte-swizzled staging buffer
    // (64 columns x 128 rows), which one thread hands to TMA as a store box: full 128-byte lines leave the SM instead of
    // 16-byte pieces of 32 different lines per warp instruction (measured: direct stores cost 15.7k clk per tile, twice
    // the tile's MMA time).
    const int q = warp & 3;
    const int r_in = q * 32 + lane;                   // row within this CTA's 128
    const uint32_t leader_acc_free = ptx::mapa(acc_free, 0);
    const uint32_t lane_addr = tmem_acc + (static_cast<uint32_t>(q * 32) << 16);
    const bool issuer = threadIdx.x == 64;
    const uint64_t pol_first = ptx::l2_policy_evict_first(), pol_last = ptx::l2_policy_evict_last();
    const bool hints = cp.l2_hints != 0;
    const uint32_t row_off = (uint32_t)r_in * 128u;
    const uint32_t sw = (uint32_t)(r_in & 7);
    const uint32_t sw64 = (uint32_t)((r_in >> 1) & 3);
    uint32_t t = 0, g = 0;                            // tiles / store boxes handled so far
    // The bias slice of a tile (256 bf16) is staged in shared memory one tile ahead: thread i fetches columns 2i, 2i+1
    // of the NEXT tile into a register before it starts on the current one and parks it in the other half of the
    // buffer afterwards, so no global-load latency (an L2 / HBM miss every time: bias slices are never reused by an SM)
    // sits inside the column loop. Measured before: 8 exposed misses per tile, epilogue 10-20k clk for 8k clk of MMAs.
    auto fetch_bias = [&](const void *D, int32_t j) -> uint32_t {
      return D ? __ldg(reinterpret_cast<const uint32_t *>(static_cast<const uint16_t *>(D) + (size_t)j * PC_BLOCK_N) + r_in) : 0u;
    };
    unsigned int epoch1 = 0;                          // the flag value of THIS launch (column-split items)
    if (cp.epoch && issuer) {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(epoch1) : "l"(cp.epoch) : "memory");
      ++epoch1;
    }
    // my rows of output tile j of layer l are complete in L2: tell whoever reads them next
    auto publish = [&](const PcItem &it, int l, int32_t j) {
      if (it.nslices == 1) {
        ptx::mbar_arrive(tile_done + 8 * j);
      } else {
        // the TMA stores of the tile have completed (bulk wait by this thread): order them before the flag, device-wide
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(cp.flags + it.flag_base + (l * PC_MAX_TILES + j) * 2 + (int)peer),
                     "r"(epoch1) : "memory");
      }
    };
    if (pair < cp.num_items) {
      const PcItem it0 = cp.items[pair];
      const uint32_t b0 = fetch_bias(cp.layers[it0.layer0].D, it0.slice);
      asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem_bias + (uint32_t)r_in * 4u), "r"(b0) : "memory");
    }
    for (int item = pair; item < cp.num_items; item += num_pairs) {
      const PcItem it = cp.items[item];
      const int32_t row0 = it.row0 + (int32_t)peer * BLOCK_M;
      for (int l = 0; l < it.num_layers; ++l) {
        const PcLayer *L = cp.layers + it.layer0 + l;
        if (issuer) tensormap_acquire(&L->tmC);
        const int32_t ln = L->n;
        const bool c64 = NARROW && L->c64 != 0;
        const int32_t xr = L->m >= BLOCK_M ? row0 % L->m : 0, xi = row0 / L->m;   // as in the producer
        const void *Dp = L->D;
        const bool relu = L->relu != 0;
        const int32_t n_tiles = L->n_tiles;
        int32_t prev_tile = -1;                         // my previous tile of this layer (its stores may still be in flight)
        for (int32_t j = it.slice; j < n_tiles; j += it.nslices, ++t) {
          const uint32_t buf = t & 1;
          // next tile's bias: same layer / next layer / first layer of this pair's next item (every slice has a tile in
          // every layer: the launcher only splits when the slice count divides all tile counts)
          uint32_t bias_next = 0;
          if (j + it.nslices < n_tiles) bias_next = fetch_bias(Dp, j + it.nslices);
          else if (l + 1 < it.num_layers) bias_next = fetch_bias(L[1].D, it.slice);
          else if (item + num_pairs < cp.num_items) {
            const PcItem nx = cp.items[item + num_pairs];
            bias_next = fetch_bias(cp.layers[nx.layer0].D, nx.slice);
          }
          ptx::mbar_wait(acc_full + 8 * buf, (t >> 1) & 1);
          ptx::tc_fence_after_sync();
          if (t < 12 && issuer) pc_stamp(cp, 4 * t + 2);
          if (l > 0 && j + it.nslices >= n_tiles && it.nslices == 1 && L[-1].d_lines > 0) {
            // this layer's last accumulator is complete: every load of my rows of the previous layer's output has been
            // consumed, nobody else reads them (CTA r only ever reads the rows it wrote) and the caller declared the
            // buffer a temporary - drop the lines from L2 before they are written back
            const PcLayer *P = L - 1;
            const int32_t lines = P->d_lines, total = lines * P->d_nrb * P->d_gk;
            const int32_t pm = P->m, rb0 = row0 / pm, rin0 = pm >= BLOCK_M ? row0 - rb0 * pm : 0;
            const char *base0 = P->d_base + (int64_t)rb0 * P->d_step_n + (int64_t)rin0 * P->d_row_bytes;
            for (int32_t idx = r_in; idx < total; idx += BLOCK_M) {
              const int32_t chunk = idx / lines, li = idx - chunk * lines, ci = chunk / P->d_gk, cj = chunk - ci * P->d_gk;
              const char *a = base0 + (int64_t)ci * P->d_step_n + (int64_t)cj * P->d_step_k + (int64_t)li * 128;
              asm volatile("discard.global.L2 [%0], 128;" ::"l"(a) : "memory");
            }
          }
          const uint32_t bias_s = smem_bias + buf * PC_BIAS_BYTES;
#pragma unroll 1
          for (int c = 0; c < PC_BLOCK_N; c += PC_OUT_COLS, ++g) {
            const uint32_t sbuf = smem_out + (g & 1) * PC_OUT_BYTES;
            // the store box issued two boxes ago has been read out of this staging buffer
            if (issuer) ptx::bulk_wait_group_read<1>();
            asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
            for (int h = 0; h < PC_OUT_COLS; h += 32) {
              uint32_t r[32];
              ptx::tmem_ld_32x32(lane_addr + buf * PC_BLOCK_N + c + h, r);
              uint32_t bw[16];
#pragma unroll
              for (int u = 0; u < 4; ++u)             // warp-uniform address: a broadcast read
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(bw[4 * u]), "=r"(bw[4 * u + 1]), "=r"(bw[4 * u + 2]), "=r"(bw[4 * u + 3])
                             : "r"(bias_s + (uint32_t)((c + h) * 2 + u * 16)));
              ptx::tmem_ld_wait();
              if (c + h == PC_BLOCK_N - 32) {         // the accumulator is in registers: hand it back to the MMA issuer
                ptx::tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive_remote(leader_acc_free + 8 * buf);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                uint32_t o[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                  float lo = __uint_as_float(r[8 * u + 2 * w]), hi = __uint_as_float(r[8 * u + 2 * w + 1]);
                  if (Dp) {
                    lo += __uint_as_float(bw[4 * u + w] << 16);
                    hi += __uint_as_float(bw[4 * u + w] & 0xffff0000u);
                  }
                  if (relu) { lo = relu_f32(lo); hi = relu_f32(hi); }
                  o[w] = pack_bf16x2(lo, hi);
                }
                const uint32_t chunk = (uint32_t)(h / 8 + u);   // 16-byte chunk of the 128-byte row
                // SWIZZLE_128B staging: [128 rows][128 bytes]; SWIZZLE_64B (32-column blocks): two [128 rows][64 bytes]
                // pieces, one per column block
                const uint32_t dst = c64 ? sbuf + (chunk >> 2) * (PC_OUT_BYTES / 2) + (uint32_t)r_in * 64u + (((chunk & 3u) ^ sw64) << 4)
                                         : sbuf + row_off + ((chunk ^ sw) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]), "r"(o[2]), "r"(o[3])
                             : "memory");
              }
            }
            ptx::fence_proxy_async();                 // my shared-memory writes -> the async proxy (TMA store)
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (issuer) {
              // a layer output that the next layer re-reads four times stays in L2; the chain's result does not
              const int32_t col = j * PC_BLOCK_N + c;
              const int32_t cn = ln >= 64 ? col % ln : 0, cj = col / ln;
              if (!hints) ptx::tma_store_4d(&L->tmC, sbuf, cn, cj, xr, xi);
              else ptx::tma_store_4d_hint(&L->tmC, sbuf, cn, cj, xr, xi, l + 1 < it.num_layers ? pol_last : pol_first);
              if (c64) {     // the second 32-column block of this 64-column group
                if (!hints) ptx::tma_store_4d(&L->tmC, sbuf + PC_OUT_BYTES / 2, 0, cj + 1, xr, xi);
                else ptx::tma_store_4d_hint(&L->tmC, sbuf + PC_OUT_BYTES / 2, 0, cj + 1, xr, xi,
                                            l + 1 < it.num_layers ? pol_last : pol_first);
              }
              ptx::bulk_commit_group();
              if (c == 0 && prev_tile >= 0 && l + 1 < it.num_layers) {
                // every store group but the one just committed is complete: my previous tile (my rows) is in L2
                ptx::bulk_wait_group<1>();
                publish(it, l, prev_tile);
              }
            }
          }
          prev_tile = j;
          // the other half of the bias buffer was last read during tile t - 1: every thread is past that
          asm volatile("st.shared.b32 [%0], %1;" ::"r"(smem_bias + (buf ^ 1u) * PC_BIAS_BYTES + (uint32_t)r_in * 4u), "r"(bias_next)
                       : "memory");
          if (t < 12 && issuer) pc_stamp(cp, 4 * t + 3);
        }
        if (l + 1 < it.num_layers && issuer && prev_tile >= 0) {
          // the layer's last tile: its stores are the only ones outstanding
          ptx::bulk_wait_group<0>();
          publish(it, l, prev_tile);
        }
      }
    }
    if (issuer) ptx::bulk_wait_group<0>();
  } else if (VNNI) {
    // ===== VNNI-2 weight converters (both CTAs): raw [k/2][n][2] rows in the slot -> swizzled MN-major tile, in place =====
    // The producer's TMA boxes deliver, per 64-column chunk, 32 k-pair rows of 256 bytes (64 columns x 2 k). Row R holds
    // exactly the bytes of the swizzled tile's rows 2R and 2R + 1 (128 bytes each), so the rewrite is in place: one
    // warp-wide 16-byte load covers two whole raw rows; lanes 2p, 2p + 1 hold columns 8p .. 8p + 3 / 8p + 4 .. 8p + 7 (both
    // k of the pair), swap halves with one shuffle, and the even lane writes the 16-byte chunk of the even k row, the odd
    // lane that of the odd k row (chunk index XOR row & 7: the SWIZZLE_128B pattern TMA would have produced for flat
    // weights). All PC_CONV_WARPS warps share every k-block (4 rows each): the rewrite sits on the ring's round trip.
    // (First version: the converters fetched the weights themselves with 16-byte global loads - 1720 clk per k-block
    // however many loads were in flight: the LSU path cannot keep as many bytes outstanding as TMA.)
    const int cw = warp - 6;                          // 0 .. PC_CONV_WARPS - 1
    // (which 8-column group a lane pair takes: within a quarter-warp the four pairs take groups {0,5,2,7} / {4,1,6,3}, so
    // that the quarter's eight 16-byte loads AND its eight 16-byte stores - two tile rows whose chunk positions differ
    // in bit 0 only - each cover all 32 banks once; the plain order 0..3 made every store a 2-way bank conflict)
    const int row_sub = lane >> 4, half = lane & 1, g8 = vnni_group_of_lane(lane);
    const uint32_t leader_full = ptx::mapa(full_bar, 0);
    constexpr int UNITS = 2 * PC_W_CHUNKS;            // 16-byte pieces per thread and k-block: 2 row groups x 2 chunks
    uint32_t q = 0;                                   // running k-block index of this CTA (all items / layers / tiles)
    for (int item = pair; item < cp.num_items; item += num_pairs) {
      const PcItem it = cp.items[item];
      for (int l = 0; l < it.num_layers; ++l) {
        const PcLayer *L = cp.layers + it.layer0 + l;
        const uint32_t kblocks = (uint32_t)(L->total_iters * ((L->n_tiles - it.slice + it.nslices - 1) / it.nslices));   // my tiles
        for (uint32_t e = 0; e < kblocks; ++e, ++q) {
          const uint32_t s = q % PC_STAGES, ph = (q / PC_STAGES) & 1u;
          if (lane == 0) ptx::mbar_wait(raw_full + 8 * s, ph);   // one poller per warp
          __syncwarp();
          if (cp.debug & 1) {
            if (lane == 0) {
              if (peer == 0) ptx::mbar_arrive(full_bar + 8 * s);
              else ptx::mbar_arrive_remote_relaxed(leader_full + 8 * s);
            }
            continue;
          }
          uint4 v[UNITS];
#pragma unroll
          for (int u = 0; u < UNITS; ++u) {
            const uint32_t R = (uint32_t)((u & 1) * 16 + cw * 2 + row_sub);     // raw row = k pair of the k-block
            const uint32_t src = smem_w + (s * PC_W_CHUNKS + (u >> 1)) * B_CHUNK_BYTES + R * 256u + (uint32_t)(2 * g8 + half) * 16u;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "r"(src));
          }
          __syncwarp();                               // every lane has read its rows before any lane overwrites them
#pragma unroll
          for (int u = 0; u < UNITS; ++u) {
            // my 4 columns x (k even | k odd) -> 4 columns of k even (lo) and 4 columns of k odd (hi)
            const uint32_t lo0 = __byte_perm(v[u].x, v[u].y, 0x5410), lo1 = __byte_perm(v[u].z, v[u].w, 0x5410);
            const uint32_t hi0 = __byte_perm(v[u].x, v[u].y, 0x7632), hi1 = __byte_perm(v[u].z, v[u].w, 0x7632);
            // the even lane keeps lo and needs its neighbour's lo; the odd lane keeps hi and needs its neighbour's hi
            const uint32_t r0 = __shfl_xor_sync(0xffffffffu, half ? lo0 : hi0, 1);
            const uint32_t r1 = __shfl_xor_sync(0xffffffffu, half ? lo1 : hi1, 1);
            const uint32_t o0 = half ? r0 : lo0, o1 = half ? r1 : lo1, o2 = half ? hi0 : r0, o3 = half ? hi1 : r1;
            const uint32_t krow = 2u * (uint32_t)((u & 1) * 16 + cw * 2 + row_sub) + (uint32_t)half;   // k row of the 64 x 64 chunk
            const uint32_t base = smem_w + (s * PC_W_CHUNKS + (u >> 1)) * B_CHUNK_BYTES;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};"
                         ::"r"(base + krow * 128u + (((uint32_t)g8 ^ (krow & 7u)) << 4)), "r"(o0), "r"(o1), "r"(o2), "r"(o3)
                         : "memory");
          }
          if (!(cp.debug & 2)) ptx::fence_proxy_async();   // my shared-memory writes -> the async proxy (the pair's MMAs)
          __syncwarp();
          if (lane == 0) {
            if (peer == 0) ptx::mbar_arrive(full_bar + 8 * s);
            else ptx::mbar_arrive_remote_relaxed(leader_full + 8 * s);
          }
        }
      }
    }
  }

  // the peer must not exit (nor free TMEM) while the leader's MMAs still read its shared memory / write its TMEM
  ptx::tc_fence_before_sync();
  __syncwarp();
  ptx::cluster_arrive();
  ptx::cluster_wait();
  if (threadIdx.x == 0) pc_stamp(cp, 61);
  if (cp.epoch && threadIdx.x == 0) {
    // the last CTA to leave closes the launch: flags of this launch (epoch + 1) become stale values for the next one
    unsigned int e;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(cp.epoch) : "memory");
    __threadfence();
    if (atomicAdd(cp.epoch + 1, 1u) == gridDim.x - 1) {
      cp.epoch[1] = 0;
      __threadfence();
      asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(cp.epoch), "r"(e + 1) : "memory");
    }
  }
  if (warp == 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc_pair(tmem_acc, 2 * PC_BLOCK_N);
  }
}


} // namespace

// ---- pair-per-chain launch ---------------------------------------------------------------------------------------------
namespace {
inline bool divides_or_multiple(int64_t v, int64_t unit) { return v > 0 && (v <= unit ? unit % v == 0 : v % unit == 0); }

// Shape rules of the kernel (per layer; a layer is a grid of tile BRGEMMs, GemmArgs::grid_*): whole 256-row work items and
// 256-column tiles; the tile dimensions must divide (or be multiples of) the box extents 128 rows / 64 k / 64 columns so
// that one TMA box is a whole number of blocks; every stride TMA sees is a multiple of 16 bytes.
bool chain_pair_supported(const KernelDesc *const *descs, const GemmArgs *args, int L, bool *vnni_out) {
  const KernelDesc &d0 = *descs[0];
  const int64_t rows = (int64_t)args[0].grid_n * d0.m;
  // whole 128-row halves: a last work item with 128 rows only runs with its second CTA on rows beyond the operands
  // (TMA fills out-of-range rows with zeros on loads and drops them on stores)
  if ((rows % BLOCK_M) != 0 || rows > (1 << 30)) return false;
  const bool vnni = (d0.gemm_flags & 2048) != 0;
  for (int l = 0; l < L; ++l) {
    const KernelDesc &d = *descs[l];
    const GemmArgs &g = args[l];
    if (!brgemm_layer_chainable(d, g)) return false;
    if (((d.gemm_flags & 2048) != 0) != vnni) return false;             // one weight layout per launch
    const int64_t n_total = (int64_t)g.grid_k * d.n, k_total = g.batch * d.k;
    if ((n_total % PC_BLOCK_N) != 0 || n_total > PC_MAX_TILES * PC_BLOCK_N) return false;
    if ((k_total % BLOCK_K) != 0 || k_total > (1 << 24)) return false;
    if (!divides_or_multiple(d.m, BLOCK_M) || !divides_or_multiple(d.k, BLOCK_K) || !divides_or_multiple(d.n, 64)) return false;
    // the innermost contiguous extent of every operand is a whole swizzle row: 128 bytes (SWIZZLE_128B) or 64 bytes
    // (SWIZZLE_64B); TMA pads anything narrower to its own line
    if (d.k < 32 || d.n < 32) return false;
    if ((d.lda % 8) != 0 || (d.ldb % 8) != 0 || (d.ldc % 8) != 0) return false;
    if (g.batch > 1 && (d.stride_a % 8) != 0) return false;
    if (g.batch > 1 && (d.stride_b % 8) != 0) return false;
    if (g.grid_n > 1 && ((g.a_step % 8) != 0 || (g.c_step_n % 8) != 0)) return false;
    if (g.grid_k > 1 && ((g.b_step % 8) != 0 || (g.c_step_k % 8) != 0)) return false;
    if (d.k < BLOCK_K && g.batch % (BLOCK_K / d.k) != 0) return false;  // a k-block is a whole number of batch elements
    if (vnni && d.vnni_factor != d0.vnni_factor) return false;
    if (vnni && d.vnni_factor == 2 && ((d.k % 2) != 0 || (d.k < BLOCK_K ? false : (d.k % BLOCK_K) != 0))) return false;
    if (vnni && d.vnni_factor != 2 && !vnni_flat_job_ok(d, g)) return false;   // VNNI-4: through a flat copy
    if (d.op == OpClass::FusedBrgemm && d.binary_kind != 0 && g.D == nullptr) return false;
    if (g.D && !aligned16(g.D)) return false;   // the epilogue reads the bias in 16-byte words
  }
  *vnni_out = vnni;
  return true;
}

// dims / strides / box of the three operand maps (see the comment above PcLayer); sizes in elements
bool encode_layer_maps(PcLayer &pl, const KernelDesc &d, const GemmArgs &g, bool vnni, const void *w_flat = nullptr) {
  const uint64_t nb = (uint64_t)g.batch, gn = (uint64_t)g.grid_n, gk = (uint64_t)g.grid_k;
  // a dimension of size 1 may carry any legal stride
  const uint64_t sa = nb > 1 ? (uint64_t)d.stride_a : (uint64_t)d.lda, sb = nb > 1 ? (uint64_t)d.stride_b : (uint64_t)d.ldb;
  const uint64_t a_step = gn > 1 ? (uint64_t)g.a_step : (uint64_t)d.lda, cn_step = gn > 1 ? (uint64_t)g.c_step_n : (uint64_t)d.ldc;
  const uint64_t b_step = gk > 1 ? (uint64_t)g.b_step : (uint64_t)d.ldb, ck_step = gk > 1 ? (uint64_t)g.c_step_k : (uint64_t)d.ldc;
  const uint32_t kx = (uint32_t)std::min<int64_t>(d.k, BLOCK_K), rx = (uint32_t)std::min<int64_t>(d.m, BLOCK_M),
                 nx = (uint32_t)std::min<int64_t>(d.n, 64);
  const bool k32 = kx == 32, n32 = nx == 32;   // 64-byte inner extents: SWIZZLE_64B, one block per box along that dimension
  pl.x64 = k32 ? 1 : 0;
  pl.w64 = (n32 && !vnni && !w_flat) ? 1 : 0;
  pl.c64 = n32 ? 1 : 0;
  pl.w_flat = w_flat ? 1 : 0;
  {
    const uint64_t dims[4] = {(uint64_t)d.k, nb, (uint64_t)d.m, gn}, str[3] = {sa, (uint64_t)d.lda, a_step};
    const uint32_t box[4] = {kx, k32 ? 1u : BLOCK_K / kx, rx, BLOCK_M / rx};
    if (!encode_map_nd(&pl.tmX, g.A, 4, dims, str, box, k32 ? 64 : 128)) return false;
  }
  if (w_flat) {
    // the flat copy: one [K][N] matrix, the ordinary 64-column x 64-k box
    const uint64_t n_total = gk * (uint64_t)d.n, k_total = nb * (uint64_t)d.k;
    const uint64_t dims[4] = {n_total, 1, k_total, 1}, str[3] = {n_total, n_total, n_total};
    const uint32_t box[4] = {64, 1, BLOCK_K, 1};
    if (!encode_map_nd(&pl.tmW, w_flat, 4, dims, str, box, 128)) return false;
  } else if (vnni) {
    // raw VNNI-2 rows: (element of the [n][2] row | column block | k pair | batch element); a box is 64 columns x 2 =
    // 256 contiguous bytes per k pair (two column blocks of 32), 32 k pairs (two batch elements when k == 32); no swizzle
    const uint32_t ex = 2 * nx, kpx = kx / 2;
    const uint64_t dims[4] = {2 * (uint64_t)d.n, gk, (uint64_t)d.k / 2, nb}, str[3] = {b_step, 2 * (uint64_t)d.ldb, sb};
    const uint32_t box[4] = {ex, 128 / ex, kpx, 32 / kpx};
    if (!encode_map_nd(&pl.tmW, g.B, 4, dims, str, box, 0)) return false;
  } else {
    // k rows of one box: all 64 of the k-block (for k == 32: both batch elements, 32 rows each)
    const uint64_t dims[4] = {(uint64_t)d.n, gk, (uint64_t)d.k, nb}, str[3] = {b_step, (uint64_t)d.ldb, sb};
    const uint32_t box[4] = {nx, n32 ? 1u : 64 / nx, kx, BLOCK_K / kx};
    if (!encode_map_nd(&pl.tmW, g.B, 4, dims, str, box, n32 ? 64 : 128)) return false;
  }
  {
    const uint64_t dims[4] = {(uint64_t)d.n, gk, (uint64_t)d.m, gn}, str[3] = {ck_step, (uint64_t)d.ldc, cn_step};
    const uint32_t box[4] = {nx, n32 ? 1u : PC_OUT_COLS / nx, rx, BLOCK_M / rx};
    if (!encode_map_nd(&pl.tmC, g.C, 4, dims, str, box, n32 ? 64 : 128)) return false;
  }
  return true;
}

template <bool VNNI, bool NARROW> bool launch_pair_kernel(const PcParams &cp, int pairs, cudaStream_t stream) {
  static std::once_flag once;
  std::call_once(once, [] {
    TPP_CUDA_CHECK(cudaFuncSetAttribute(mlp_chain_pair_kernel<VNNI, NARROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, PC_SMEM));
  });
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(2 * pairs));
  cfg.blockDim = dim3(VNNI ? PC_THREADS_VNNI : NUM_THREADS);
  cfg.dynamicSmemBytes = PC_SMEM;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[3];
  attrs[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[0].val.programmaticStreamSerializationAllowed = 1;
  attrs[1].id = cudaLaunchAttributeClusterDimension;
  attrs[1].val.clusterDim.x = 2;
  attrs[1].val.clusterDim.y = 1;
  attrs[1].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  // column-split items wait for each other's tiles through flags in global memory: every pair must be resident
  if (cp.flags && !prepare_resident_launch(reinterpret_cast<const void *>(mlp_chain_pair_kernel<VNNI, NARROW>), &cfg, attrs))
    return false;
  TPP_CUDA_CHECK(cudaLaunchKernelEx(&cfg, mlp_chain_pair_kernel<VNNI, NARROW>, cp));
  return true;
}
}  // namespace


// Launch a prefix of chains [0, num_chains) as ONE launch of mlp_chain_pair_kernel: every chain is cut into blocks of
// 256 batch rows, every block is a work item of one CTA pair. Without `force` it is taken only when the launch carries
// enough items to occupy a useful share of the 74 pairs (a single pair needs ~50 us for a 3 x 1024^2 chain; the pass
// kernels finish a lone flat chain in ~11 us). Returns the number of chains launched (0: not applicable).
int launch_brgemm_chains_pair(const KernelDesc *const *descs, const GemmArgs *args, const int *first, const int *len,
                              int num_chains, cudaStream_t stream, bool force) {
  static const int min_items = [] {
    const char *e = getenv("TPP_XSMM_CHAIN_PAIR_MIN");   // 0 disables the kernel
    return e ? atoi(e) : 12;
  }();
  if (min_items <= 0 || num_chains < 1) return 0;
  int take = 0;
  int64_t items = 0, layers = 0;
  bool vnni = false;
  {
    std::vector<ByteRange> in_all, out_all;
    while (take < num_chains) {
      const int c = take;
      bool v = false;
      if (!chain_pair_supported(descs + first[c], args + first[c], len[c], &v)) break;
      if (take > 0 && v != vnni) break;   // one instantiation per launch
      vnni = v;
      std::vector<ByteRange> in, out;
      chain_ranges(descs + first[c], args + first[c], len[c], in, out);
      bool indep = true;
      for (const ByteRange &o : out) {
        for (const ByteRange &x : in_all) indep = indep && !overlaps(o, x);
        for (const ByteRange &x : out_all) indep = indep && !overlaps(o, x);
      }
      for (const ByteRange &i : in)
        for (const ByteRange &x : out_all) indep = indep && !overlaps(i, x);
      if (!indep) break;
      in_all.insert(in_all.end(), in.begin(), in.end());
      out_all.insert(out_all.end(), out.begin(), out.end());
      items += ((int64_t)args[first[c]].grid_n * descs[first[c]]->m + PC_ROWS - 1) / PC_ROWS;
      layers += len[c];
      ++take;
    }
  }
  if (take == 0 || (!force && items < min_items)) return 0;
  static const int max_pairs = [] {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const char *e = getenv("TPP_XSMM_CHAIN_PAIRS");
    const int p = e ? atoi(e) : sms / 2;
    return p < 1 ? 1 : p;
  }();
  // Few row blocks (a lone forward pass is ONE): deal the output tiles of every layer out to `nslices` pairs per row
  // block, the largest power of two that divides every layer's tile count and still fits one round of resident pairs
  // (the slices wait for each other, so they must all be on the machine at once).
  static const int split_max = [] { const char *e = getenv("TPP_XSMM_PAIR_SPLIT"); return e ? atoi(e) : 16; }();
  constexpr int kFlagsPerRowBlock = CHAIN_MAX_LAYERS * PC_MAX_TILES * 2;
  int nslices = 1;
  for (int d = 16; d >= 2; d /= 2) {
    if (d > split_max || items * d > max_pairs) continue;
    bool ok = true;
    for (int c = 0; c < take && ok; ++c)
      for (int l = 0; l < len[c] && ok; ++l) {
        const int64_t nt = (int64_t)args[first[c] + l].grid_k * descs[first[c] + l]->n / PC_BLOCK_N;
        ok = (nt % d) == 0;
      }
    if (ok) { nslices = d; break; }
  }
  const int64_t row_blocks = items;
  items *= nslices;
  std::vector<PcLayer> hl((size_t)layers);
  std::vector<PcItem> hi((size_t)items);
  size_t nl = 0, ni = 0;
  int32_t row_block = 0;
  bool grids = false;
  static const bool force_narrow = [] { const char *e = getenv("TPP_XSMM_PAIR_NARROW"); return e && e[0] == '1'; }();   // A/B
  bool narrow = force_narrow;
  // VNNI-4 weights (mlir-gen --vnni=4): the converter warps rewrite factor 2 only, so every distinct weight buffer of the
  // launch gets a flat [K][N] copy made by one kernel in front of this one (vnni_flat.cu; graph-owned scratch, rebuilt by
  // every replay) and the layers read the copies with the flat-weight instantiation
  // VNNI-2 weights that MANY work items of the launch share (a served model: one set of weights for all the input
  // batches of the launch, or the row blocks of a large batch) take the flat copy too: the in-place rewrite costs every
  // item 32 KiB of shared-memory traffic per k-block (982 against 1285 TF/s with L2-resident weights), the copy costs the
  // launch 4 bytes per weight element once
  const int vfactor = vnni ? descs[first[0]]->vnni_factor : 0;
  bool w_flat = vnni && vfactor != 2;
  if (vnni && vfactor == 2) {
    static const int min_reuse = [] { const char *e = getenv("TPP_XSMM_PAIR_FLATW_REUSE"); return e ? atoi(e) : 4; }();
    std::vector<const void *> distinct;
    int64_t uses = 0;
    bool ok = min_reuse > 0;
    for (int c = 0; c < take && ok; ++c) {
      const int64_t blocks = ((int64_t)args[first[c]].grid_n * descs[first[c]]->m + PC_ROWS - 1) / PC_ROWS;
      for (int l = 0; l < len[c] && ok; ++l) {
        const GemmArgs &g = args[first[c] + l];
        ok = vnni_flat_job_ok(*descs[first[c] + l], g);
        uses += blocks;
        if (std::find(distinct.begin(), distinct.end(), g.B) == distinct.end()) distinct.push_back(g.B);
      }
    }
    w_flat = ok && !distinct.empty() && uses >= (int64_t)min_reuse * (int64_t)distinct.size();
  }
  std::vector<VnniFlatJob> wf;
  auto flat_copy_of = [&](const KernelDesc &d, const GemmArgs &g) -> const void * {
    for (const VnniFlatJob &e : wf)
      if (e.src == g.B) return e.dst;
    void *buf = nullptr;
    TPP_CUDA_CHECK(cudaMalloc(&buf, (size_t)g.batch * d.k * g.grid_k * d.n * sizeof(uint16_t)));
    capture_adopt(buf);
    wf.push_back(vnni_flat_job(d, g, buf));
    return buf;
  };
  if (w_flat) vnni = false;   // kernel instantiation: flat weights
  for (int c = 0; c < take; ++c) {
    const int32_t layer0 = (int32_t)nl;
    for (int l = 0; l < len[c]; ++l) {
      const KernelDesc &d = *descs[first[c] + l];
      const GemmArgs &g = args[first[c] + l];
      PcLayer &pl = hl[nl++];
      memset(&pl, 0, sizeof(pl));
      if (!encode_layer_maps(pl, d, g, vnni, w_flat ? flat_copy_of(d, g) : nullptr)) return 0;
      static const bool dbg = getenv("TPP_XSMM_DEBUG") != nullptr;
      if (dbg)
        fprintf(stderr, "pair-chain layer: chain %d layer %d tile m=%lld n=%lld k=%lld batch=%lld lda=%lld ldb=%lld ldc=%lld sa=%lld "
                        "sb=%lld grid %d x %d steps a=%lld b=%lld cn=%lld ck=%lld d=%lld vnni=%d A=%p B=%p C=%p D=%p\n", c, l,
                (long long)d.m, (long long)d.n, (long long)d.k, (long long)g.batch, (long long)d.lda, (long long)d.ldb,
                (long long)d.ldc, (long long)d.stride_a, (long long)d.stride_b, g.grid_n, g.grid_k, (long long)g.a_step,
                (long long)g.b_step, (long long)g.c_step_n, (long long)g.c_step_k, (long long)g.d_step, (int)vnni, g.A, g.B, g.C,
                g.D);
      grids = grids || g.is_grid();
      narrow = narrow || pl.x64 || pl.w64 || pl.c64;
      pl.D = (d.op == OpClass::FusedBrgemm && g.D && d.binary_kind == 1) ? g.D : nullptr;
      pl.W = g.B;
      pl.w_col_step = g.grid_k > 1 ? g.b_step : 0;
      pl.w_batch_step = g.batch > 1 ? d.stride_b : 0;
      pl.w_ldb = d.ldb;
      pl.m = (int32_t)d.m; pl.n = (int32_t)d.n; pl.k = (int32_t)d.k;
      pl.k_bstep = d.k < BLOCK_K ? (int32_t)(BLOCK_K / d.k) : 1;
      pl.total_iters = (int32_t)(g.batch * d.k / BLOCK_K);
      pl.n_tiles = (int32_t)((int64_t)g.grid_k * d.n / PC_BLOCK_N);
      pl.relu = (d.op == OpClass::FusedBrgemm && d.unary_kind == 5) ? 1 : 0;
      pl.vnni = vnni ? 1 : 0;
      // a temporary the next layer of the chain consumes: geometry of one CTA's 128 rows as contiguous chunks of lines
      pl.d_lines = 0;
      static const bool discard_off = [] { const char *e = getenv("TPP_XSMM_DISCARD"); return e && e[0] == '0'; }();
      if (!discard_off && l + 1 < len[c] && nslices == 1 && ((int64_t)g.grid_n * d.m) % PC_ROWS == 0) {
        const LayerRanges lr = layer_ranges(d, g);
        const int64_t rows_blk = std::min<int64_t>(d.m, BLOCK_M);          // my rows inside one row block
        const int64_t chunk_bytes = rows_blk * d.n * 2;                      // contiguous when the block's rows are packed
        const bool packed = d.ldc == d.n;
        const int64_t step_n = (g.grid_n > 1 ? g.c_step_n : 0) * 2, step_k = (g.grid_k > 1 ? g.c_step_k : 0) * 2;
        if (packed && range_is_temporary(lr.c.lo, (size_t)(lr.c.hi - lr.c.lo)) && (reinterpret_cast<uintptr_t>(g.C) % 128) == 0 &&
            (chunk_bytes % 128) == 0 && (step_n % 128) == 0 && (step_k % 128) == 0 && ((int64_t)BLOCK_M * d.ldc * 2) % 128 == 0 &&
            chunk_bytes / 128 < (1 << 20)) {
          pl.d_base = static_cast<const char *>(g.C);
          pl.d_step_n = step_n;
          pl.d_step_k = step_k;
          pl.d_row_bytes = d.ldc * 2;
          pl.d_lines = (int32_t)(chunk_bytes / 128);
          pl.d_nrb = (int32_t)(BLOCK_M / rows_blk);
          pl.d_gk = g.grid_k;
        }
      }
    }
    const int64_t rows = (int64_t)args[first[c]].grid_n * descs[first[c]]->m;
    for (int64_t r = 0; r < rows; r += PC_ROWS, ++row_block) {
      for (int sl = 0; sl < nslices; ++sl) {      // the slices of a row block sit on neighbouring pairs (shared input rows)
        PcItem &pi = hi[ni++];
        memset(&pi, 0, sizeof(pi));
        pi.layer0 = layer0;
        pi.num_layers = len[c];
        pi.row0 = (int32_t)r;
        pi.slice = sl;
        pi.nslices = nslices;
        pi.flag_base = row_block * kFlagsPerRowBlock;
      }
    }
  }
  // the table is written now (not captured): a graph replay only launches the kernel that reads it
  const size_t lbytes = hl.size() * sizeof(PcLayer), ibytes = (hi.size() * sizeof(PcItem) + 127) & ~(size_t)127;
  char *table = nullptr;
  TPP_CUDA_CHECK(cudaMalloc(&table, lbytes + ibytes));
  TPP_CUDA_CHECK(cudaMemcpyAsync(table, hl.data(), lbytes, cudaMemcpyHostToDevice, table_stream()));
  TPP_CUDA_CHECK(cudaMemcpyAsync(table + lbytes, hi.data(), hi.size() * sizeof(PcItem), cudaMemcpyHostToDevice, table_stream()));
  TPP_CUDA_CHECK(cudaStreamSynchronize(table_stream()));
  capture_adopt(table);
  PcParams cp;
  cp.layers = reinterpret_cast<const PcLayer *>(table);
  cp.items = reinterpret_cast<const PcItem *>(table + lbytes);
  cp.num_items = (int32_t)items;
  cp.flags = nullptr;
  cp.epoch = nullptr;
  if (nslices > 1) {
    const size_t words = (size_t)row_blocks * kFlagsPerRowBlock + 2;
    cp.flags = static_cast<unsigned int *>(capture_owned_zeroed(words * sizeof(unsigned int)));
    cp.epoch = cp.flags + (size_t)row_blocks * kFlagsPerRowBlock;
  }
  static const bool hints_on = [] { const char *e = getenv("TPP_XSMM_CHAIN_PAIR_HINTS"); return !(e && e[0] == '0'); }();
  cp.l2_hints = hints_on ? 1 : 0;
  static const int debug = [] { const char *e = getenv("TPP_XSMM_PAIR_DEBUG"); return e ? atoi(e) : 0; }();
  cp.debug = debug;
  cp.trace = nullptr;
  static const bool pc_trace_on = [] { const char *e = getenv("TPP_XSMM_TC_TRACE"); return e && atoi(e) == 4; }();
  if (pc_trace_on) {
    if (!g_pc_trace) {
      TPP_CUDA_CHECK(cudaMalloc(&g_pc_trace, sizeof(unsigned long long) * 2 * 148 * PC_TRACE_SLOTS));
      TPP_CUDA_CHECK(cudaMemset(g_pc_trace, 0, sizeof(unsigned long long) * 2 * 148 * PC_TRACE_SLOTS));
    }
    cp.trace = g_pc_trace;
  }
  // balanced: the fewest pairs that still need the minimal number of rounds
  const int rounds = (int)((items + max_pairs - 1) / max_pairs);
  const int pairs = (int)((items + rounds - 1) / rounds);
  g_pc_trace_ctas = 2 * pairs;
  if (w_flat) launch_vnni_weights_to_flat(wf.data(), (int)wf.size(), stream);
  const bool launched = vnni && narrow ? launch_pair_kernel<true, true>(cp, pairs, stream)
                        : vnni         ? launch_pair_kernel<true, false>(cp, pairs, stream)
                        : narrow       ? launch_pair_kernel<false, true>(cp, pairs, stream)
                                       : launch_pair_kernel<false, false>(cp, pairs, stream);
  if (!launched) return 0;
  char split_tag[16] = "";
  if (nslices > 1) snprintf(split_tag, sizeof(split_tag), "_split%d", nslices);
  set_last_name("mlp_chain_bf16_%dx%dlayers_pair256x256%s%s%s", (int)row_blocks, len[0], grids ? "_blocked" : "",
                vfactor == 2 ? "_vnni2" : vfactor == 4 ? "_vnni4" : "", split_tag);
  return take;
}


} // namespace tpp
