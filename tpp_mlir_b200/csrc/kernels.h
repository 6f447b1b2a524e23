// kernels.h - host-side launch entry points of the sm_100a kernels.
// Every pointer here is a DEVICE pointer (already offset to the first element).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <vector>

#include "kernel_desc.h"

namespace tpp {

// broadcast mode of one eltwise operand
enum : int { kBcastNone = 0, kBcastRow = 1, kBcastCol = 2, kBcastScalar = 3, kBcastImm = 4 /* scalar passed by value */ };
// eltwise opcode
enum : int { kOpIdentity = 0, kOpZero = 1, kOpRelu = 2, kOpAdd = 3, kOpMul = 4, kOpSub = 5, kOpDiv = 6 };

struct EltwiseArgs {
  const void *in0 = nullptr;
  const void *in1 = nullptr;
  void *out = nullptr;
  int64_t m = 0, n = 0, ld0 = 0, ld1 = 0, ldo = 0;
  int mode0 = 0, mode1 = 0, op = 0;
  int64_t dtype = 0;
  float imm = 0.f; // operand 0 when mode0 == kBcastImm (xsmm_unary_scalar_invoke)
};

// unary identity/zero/relu and binary add/mul/sub/div, with broadcasts
void launch_eltwise(const EltwiseArgs &a, cudaStream_t stream);
// out[j*ldo+i] = in[i*ldi+j], i<m, j<n (bit copy; es = element size 2 or 4)
void launch_transpose(const void *in, void *out, int64_t m, int64_t n, int64_t ldi, int64_t ldo, int es,
                      cudaStream_t stream);
// bf16 [K=m][N=n] (ldi) -> [K/2][N][2] (ldo in pairs), and the inverse
void launch_vnni2_pack(const void *in, void *out, int64_t m, int64_t n, int64_t ldi, int64_t ldo,
                       cudaStream_t stream);
void launch_vnni2_unpack(const void *in, void *out, int64_t m, int64_t n, int64_t ldi, int64_t ldo,
                         cudaStream_t stream);
// bf16 [K=m][N=n] (ldi) -> [K/4][N][4] (ldo in quads), and the inverse
void launch_vnni4_pack(const void *in, void *out, int64_t m, int64_t n, int64_t ldi, int64_t ldo, cudaStream_t stream);
void launch_vnni4_unpack(const void *in, void *out, int64_t m, int64_t n, int64_t ldi, int64_t ldo, cudaStream_t stream);

// one tile of a batched tile move (identity copy or transpose, same m / n / ldi / ldo for the whole batch)
struct TilePtrs {
  const void *in;
  void *out;
};
// num_tiles moves in ONE launch; dev_tiles is a device array. vec16_ok: every pointer and both row pitches are
// multiples of 16 bytes and so is a row of n elements (identity only)
void launch_tile_batch(const TilePtrs *dev_tiles, int64_t num_tiles, bool transpose, int64_t m, int64_t n, int64_t ldi,
                       int64_t ldo, int es, bool vec16_ok, cudaStream_t stream);

// a run of tile copies that walks a regular grid - tile t = i * J + j reads in0 + i * in_outer + j * in_inner and writes
// out0 + i * out_outer + j * out_inner (byte steps) - as ONE TMA-to-TMA copy kernel (tile_grid.cu); false = TMA cannot
// express the run (nothing launched)
bool launch_tile_grid(const void *in0, void *out0, int64_t J, int64_t I, int64_t in_inner, int64_t in_outer, int64_t out_inner,
                      int64_t out_outer, int64_t m, int64_t n, int64_t ldi, int64_t ldo, int es, cudaStream_t stream);

// [p, p + bytes) lies inside a range the caller marked with xsmm_cuda_mark_temporary (runtime.cu)
bool range_is_temporary(const void *p, size_t bytes);

struct GemmArgs {
  const void *A = nullptr;
  const void *B = nullptr;
  void *C = nullptr;
  const void *D = nullptr; // bias vector (fused add, bcast_col_in0) or nullptr
  int64_t batch = 1;
  // true when B (and D) were not written by any of the last few kernels this thread launched: the
  // kernel may then fetch B before the programmatic-dependent-launch wait (weights of an MLP layer)
  bool b_independent = false;
  // same for A: the fused chain kernel may then fetch its first layer's activations (and run that layer's MMAs,
  // which only touch shared / tensor memory) before the wait; stores always come after it
  bool a_independent = false;
  // programmatic dependent launch allowed. The runtime clears it once per window of kPdlWindow PDL launches: a launch
  // in plain stream order waits for EVERYTHING before it, so at most one window of kernels can ever be co-resident and
  // the "not written by the last kPdlWindow kernels" test behind b_independent / a_independent is a proof, not a hope
  bool pdl = true;
  // B is VNNI-2 packed ([batch][k/2][ldb][2]) although the descriptor passed to launch_brgemm_tc is the flat twin: the
  // CTA-pair kernel converts it in shared memory (launch_brgemm_tc returns false when it would pick another kernel)
  bool b_vnni2 = false;
  // A LAYER regrouped from the invokes recorded during graph capture: grid_n x grid_k invokes of ONE descriptor, tile
  // (i, j) being the invoke on A + i a_step, B + j b_step, C + i c_step_n + j c_step_k, D + j d_step (steps in elements).
  // This is what the reference's tiled loop nest emits per layer (block-packed operands, one
  // BRGEMM per (iN, iK) output block); 1 x 1 is a plain invoke. Only the pair-per-chain kernel takes grids.
  int32_t grid_n = 1, grid_k = 1;
  int64_t a_step = 0, b_step = 0, c_step_n = 0, c_step_k = 0, d_step = 0;
  bool is_grid() const { return grid_n != 1 || grid_k != 1; }
};
constexpr int kPdlWindow = 32;

// generic FFMA BRGEMM (any dtype/ld/stride, VNNI-B), fused epilogue
void launch_brgemm_simt(const KernelDesc &d, const GemmArgs &g, cudaStream_t stream);

// tcgen05 BRGEMM. Returns false (and launches nothing) if the operands of THIS
// invoke are not TMA-compatible (pointer alignment); the caller then uses the
// generic kernel.
bool brgemm_tc_supported(const KernelDesc &d);
void brgemm_tc_configure(KernelDesc &d);
bool launch_brgemm_tc(const KernelDesc &d, const GemmArgs &g, cudaStream_t stream);
// L consecutive layers (C of one is A of the next) in one persistent kernel: see brgemm_tc.cu
bool brgemm_chain_linked(const KernelDesc *const *descs, const GemmArgs *args, int L);      // the layers form a chain
bool brgemm_layer_chainable(const KernelDesc &d, const GemmArgs &g);   // a layer a chain kernel could take (bf16, beta_0, ...)
bool brgemm_chain_supported(const KernelDesc *const *descs, const GemmArgs *args, int L);   // ... the pass kernels can run
bool launch_brgemm_chain(const KernelDesc *const *descs, const GemmArgs *args, int L, cudaStream_t stream);
// several chains (chain c = layers [first[c], first[c] + len[c]), each accepted by brgemm_chain_supported) as ONE launch
// of the feature-major chain kernel, pairs of mutually independent chains interleaved; returns the number of chains
// launched (a prefix), 0 if the kernel does not apply
int launch_brgemm_chains_ft(const KernelDesc *const *descs, const GemmArgs *args, const int *first, const int *len,
                            int num_chains, cudaStream_t stream);
// same contract, for launches that carry many independent chains: one CTA pair (cta_group::2, 256 x 256 tiles) walks a
// whole chain for a block of 256 batch rows; taken when the prefix has at least TPP_XSMM_CHAIN_PAIR_MIN (12) such blocks
// `force`: take the prefix whatever its size (layers that are grids of small tile invokes or have VNNI-2 weights have no
// other tensor-core kernel: the alternative is one launch per tile)
int launch_brgemm_chains_pair(const KernelDesc *const *descs, const GemmArgs *args, const int *first, const int *len,
                              int num_chains, cudaStream_t stream, bool force = false);
// flat [K][N] copies of VNNI-packed (factor 2 or 4) weights, made by one kernel in front of a chain kernel (vnni_flat.cu)
struct VnniFlatJob {
  const void *src;   // [gk column blocks][nb batch elements][k / v][ldb][v]
  void *dst;         // [nb * k][gk * n]
  int64_t ldb, stride_b, b_step;
  int32_t n, k, nb, gk, v;
};
bool vnni_flat_job_ok(const KernelDesc &d, const GemmArgs &g);
VnniFlatJob vnni_flat_job(const KernelDesc &d, const GemmArgs &g, void *dst);
void launch_vnni_weights_to_flat(const VnniFlatJob *jobs, int count, cudaStream_t stream);
// device tables allocated by chain launches since the last call (owned by the graph being captured)
void brgemm_tc_take_capture_allocs(std::vector<void *> &out);
const char *brgemm_tc_last_name();   // tile configuration of this thread's last tcgen05 launch
int brgemm_tc_take_extra_launches();   // helper kernels the last chain launches put in front of themselves (since the last call)
void brgemm_tc_dump_trace();   // debug, TPP_XSMM_TC_TRACE=2

} // namespace tpp
