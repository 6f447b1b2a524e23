// tpp_run_standin.cpp - what `mlir-gen --kernel=const --bias --relu --float-type=bf16 ... | tpp-run -n N` executes,
// written by hand (LLVM / MLIR are not available in this image, so tpp-run itself cannot be built).
//
// The JIT-compiled main() of tpp-run (lib/TPP/Runner/MLIRBench.cpp:208-300, TppRunnerWrapper.cpp:89-131) allocates the
// kernel arguments as host globals, fills them (TensorInit), hoists the xsmm dispatches, runs clamp(N/100,1,50)
// warm-up calls and N timed calls of the kernel between perf_start_timer / perf_stop_timer, and prints the mean. The
// kernel is the loop nest of SURVEY.md Appendix B: per layer one xsmm_fused_brgemm_invoke per (iN, iK) output block
// (plus the AMX tile-config invokes the bf16 pipeline wraps around them). This program does exactly that and calls
// ONLY the C-ABI of include/tpp_xsmm_abi.h.
//
//   --mode strict   plain host pointers, nothing registered: an UNMODIFIED tpp-run linked against this library
//   --mode device   arguments live on the device (what MLIRBench::registerOnGpu gives the patched runner,
//                   patches/0004): invokes take the in-place path, one launch per invoke
//   --mode lazy     device arguments + xsmm_cuda_set_lazy(1): the same invoke loop, no graph calls; invokes are queued and
//                   launched fused at perf_stop_timer (what TPP_XSMM_LAZY=1 gives the patched runner WITHOUT patches/0005)
//   --mode graph    device arguments + the timed body recorded once and replayed (patches/0005)
//
//   tpp_run_standin [--batch 256] [--layers 1024,1024,1024,1024] [--tiles 32,32,32] [--vnni 2] [-n 100] [--seed 123]
//                   [--mode strict|device|lazy|graph]
// prints: seconds per iteration (mean), GFLOP/s by mlir-gen's flop count, and a checksum of the output.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "tpp_xsmm_abi.h"

namespace {
uint16_t f32_to_bf16(float f) {   // round to nearest even (mlir Float16bits.h)
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
std::vector<int64_t> ints(const char *s) {
  std::vector<int64_t> v;
  for (const char *p = s; *p;) {
    v.push_back(strtoll(p, const_cast<char **>(&p), 10));
    if (*p == ',') ++p;
  }
  return v;
}
// tpp-run --init-type normal: clamp(N(0, 0.2), 0, 1) from one std::default_random_engine per program, RNE to bf16
// (include/TPP/Transforms/Utils/TensorInitFloat.h:133-142)
struct NormalInit {
  std::default_random_engine gen;
  std::normal_distribution<float> dist{0.0f, 0.2f};
  explicit NormalInit(int seed) : gen((unsigned)seed) {}
  void fill(uint16_t *p, size_t n) {
    for (size_t i = 0; i < n; ++i) p[i] = f32_to_bf16(std::min(1.0f, std::max(0.0f, dist(gen))));
  }
};
uint16_t *host_alloc(size_t elems) {   // memref globals are 128-byte aligned (BuilderUtils.cpp createDenseMemref)
  void *p = nullptr;
  if (posix_memalign(&p, 128, elems * 2 + 128) != 0) exit(1);
  memset(p, 0, elems * 2);
  return static_cast<uint16_t *>(p);
}
}  // namespace

int main(int argc, char **argv) {
  int64_t batch = 256, n_iter = 100, vnni = 2, seed = 123;
  std::vector<int64_t> layers = {1024, 1024, 1024, 1024}, tiles = {32, 32, 32};
  std::string mode = "strict";
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string a = argv[i];
    if (a == "--batch") batch = atoll(argv[i + 1]);
    else if (a == "--layers") layers = ints(argv[i + 1]);
    else if (a == "--tiles") tiles = ints(argv[i + 1]);
    else if (a == "--vnni") vnni = atoll(argv[i + 1]);
    else if (a == "-n") n_iter = atoll(argv[i + 1]);
    else if (a == "--seed") seed = atoll(argv[i + 1]);
    else if (a == "--mode") mode = argv[i + 1];
    else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
  }
  if (tiles.size() != 3 || layers.size() < 2 || (vnni != 0 && vnni != 2)) { fprintf(stderr, "bad shape options\n"); return 2; }
  const int64_t bn = tiles[0], bk = tiles[1], bc = tiles[2];
  const int64_t L = (int64_t)layers.size() - 1;
  if (batch % bn) { fprintf(stderr, "batch must be a multiple of the bn tile\n"); return 2; }
  for (int64_t l = 0; l < L; ++l)
    if (layers[l] % bc || layers[l + 1] % bk) { fprintf(stderr, "layer sizes must be multiples of the tiles\n"); return 2; }

  // ---- arguments: splat constants first in op order (W1, b1, W2, b2, ...), then the kernel argument (MLIRBench.cpp:111-164)
  NormalInit init((int)seed);
  std::vector<uint16_t *> W(L), B(L), act(L + 1);
  for (int64_t l = 0; l < L; ++l) {
    W[l] = host_alloc(layers[l] * layers[l + 1]);   // block-packed [K/bk][C/bc][bc][bk] (VNNI: [..][bc/2][bk][2]): same element count
    B[l] = host_alloc(layers[l + 1]);
    init.fill(W[l], layers[l] * layers[l + 1]);
    init.fill(B[l], layers[l + 1]);
  }
  for (int64_t l = 0; l <= L; ++l) act[l] = host_alloc(batch * layers[l]);
  init.fill(act[0], batch * layers[0]);

  // ---- dispatch, hoisted out of the loops (one handle: every layer has the same tile shape)
  const int64_t gflags = XSMM_GEMM_FLAG_BETA_0 | (vnni ? XSMM_GEMM_FLAG_ROWMAJOR_B_VNNI : 0) | 64 | 128;
  const int64_t h = xsmm_fused_brgemm_dispatch(2, bn, bk, bc, bc, bk, bk, bn * bc, bc * bk, gflags, 0, XSMM_UNARY_RELU,
                                               XSMM_BINARY_FLAG_BCAST_COL_IN_0, XSMM_BINARY_ADD);
  const int64_t tc = xsmm_intel_amx_tile_config_dispatch(2, bn, bk, bc, bc, bk, bk, bn * bc, bc * bk, gflags);

  // ---- residency
  std::vector<uint16_t *> dW = W, dB = B, dact = act;
  const bool device = mode != "strict";
  if (device) {   // gpu.alloc + gpu.memcpy of every argument (MLIRBench::registerOnGpu)
    for (int64_t l = 0; l < L; ++l) {
      xsmm_cuda_register_host(W[l], layers[l] * layers[l + 1] * 2, 1);
      xsmm_cuda_register_host(B[l], layers[l + 1] * 2, 1);
      dW[l] = static_cast<uint16_t *>(xsmm_cuda_device_ptr(W[l]));
      dB[l] = static_cast<uint16_t *>(xsmm_cuda_device_ptr(B[l]));
    }
    for (int64_t l = 0; l <= L; ++l) {
      xsmm_cuda_register_host(act[l], batch * layers[l] * 2, 1);
      dact[l] = static_cast<uint16_t *>(xsmm_cuda_device_ptr(act[l]));
      // the outputs of all layers but the last are buffers the generated kernel allocates and frees itself
      // (patches/0006 marks them)
      if (l > 0 && l < L) xsmm_cuda_mark_temporary(dact[l], batch * layers[l] * 2);
    }
  }

  if (mode == "lazy") xsmm_cuda_set_lazy(1);
  char amx_state[64];
  auto kernel = [&]() {   // the loop nest the pipeline emits (scf.parallel serialised on one stream)
    for (int64_t l = 0; l < L; ++l) {
      const int64_t nb_c = layers[l] / bc, nb_k = layers[l + 1] / bk;
      for (int64_t in = 0; in < batch / bn; ++in)
        for (int64_t ik = 0; ik < nb_k; ++ik) {
          xsmm_intel_amx_tile_config_invoke(2, tc, amx_state, 0);
          xsmm_fused_brgemm_invoke(2, h, dact[l], in * nb_c * bn * bc, dW[l], ik * nb_c * bc * bk, dact[l + 1],
                                   (in * nb_k + ik) * bn * bk, dB[l], ik * bk, nb_c);
          xsmm_intel_amx_tile_config_invoke(2, tc, amx_state, 0);
        }
    }
  };
  int64_t graph = 0;
  if (mode == "graph") {
    if (xsmm_cuda_graph_begin() != 0) return 1;
    kernel();
    graph = xsmm_cuda_graph_end();
    if (!graph) return 1;
  }
  auto iteration = [&]() {
    if (graph) xsmm_cuda_graph_launch(graph);
    else kernel();
  };
  const int64_t warm = std::min<int64_t>(std::max<int64_t>(n_iter / 100, 1), 50);
  for (int64_t i = 0; i < warm; ++i) iteration();
  const int64_t t0 = perf_start_timer();
  for (int64_t i = 0; i < n_iter; ++i) iteration();
  const double mean = perf_stop_timer(t0) / (double)n_iter;

  if (device) {
    xsmm_cuda_update_host(act[L], batch * layers[L] * 2);
    xsmm_cuda_sync();
  }
  double checksum = 0.0;
  for (int64_t i = 0; i < batch * layers[L]; ++i) checksum += bf16_to_f32(act[L][i]);
  double flops = 0.0;   // mlir-gen's BENCH_TOTAL_FLOPS (tools/mlir-gen/MLIRGen.cpp:313-334)
  for (int64_t l = 0; l < L; ++l) flops += 2.0 * batch * layers[l] * layers[l + 1] + 2.0 * batch * layers[l + 1];
  printf("{\"mode\": \"%s\", \"seconds_per_iteration\": %.9f, \"gflops\": %.3f, \"iterations\": %lld, \"kernel\": \"%s\", "
         "\"launches\": %lld, \"checksum\": %.6f}\n",
         mode.c_str(), mean, flops / mean / 1e9, (long long)n_iter, xsmm_cuda_last_kernel(), (long long)xsmm_cuda_launch_count(),
         checksum);
  if (graph) xsmm_cuda_graph_destroy(graph);
  return 0;
}
