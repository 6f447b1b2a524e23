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
//   tpp_run_standin [--batch 256] [--layers 1024,1024,1024,1024] [--tiles 32,32,32] [--vnni 0|2|4] [-n 100] [--seed 123]
//                   [--float-type bf16|f32] [--bias 0|1] [--relu 0|1] [--kernel const|args]
//                   [--mode strict|device|lazy|graph]
//   tpp_run_standin --mlir-gen "--kernel=args --bias --relu --float-type=bf16 --vnni=4 --batch=256 --layers=768,768
//                   --tiles=64,64,64" [-n 100] [--mode ...]
//       the flags of a reference benchmark line as they stand in benchmarks/config/*/*.json, read the way mlir-gen reads
//       them (tools/mlir-gen/mlir-gen.cpp: no --bias / --relu = a plain matmul -> xsmm_brgemm_*, --float-type defaults to
//       f32, --kernel to const; without --tiles the pipeline's default 32 x 32 x 32 packing; a bf16 kernel without --vnni
//       gets the VNNI factor the library reports, 2). --kernel=args: every layer's output is a kernel argument, so nothing
//       is marked temporary. --vnni=4: the factor libxsmm_cpuid_dot_pack_factor has to answer is set through
//       TPP_XSMM_VNNI before the dispatch.
// prints: seconds per iteration (mean), GFLOP/s by mlir-gen's flop count, and a checksum of the output.
#include <algorithm>
#include <cctype>
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
  void fill(float *p, size_t n) {
    for (size_t i = 0; i < n; ++i) p[i] = std::min(1.0f, std::max(0.0f, dist(gen)));
  }
  void fill(void *p, size_t n, bool bf16) {
    if (bf16) fill(static_cast<uint16_t *>(p), n);
    else fill(static_cast<float *>(p), n);
  }
};
void *host_alloc(size_t elems, size_t es) {   // memref globals are 128-byte aligned (BuilderUtils.cpp createDenseMemref)
  void *p = nullptr;
  if (posix_memalign(&p, 128, elems * es + 128) != 0) exit(1);
  memset(p, 0, elems * es);
  return p;
}

struct Options {
  int64_t batch = 256, n_iter = 100, vnni = 2, seed = 123;
  std::vector<int64_t> layers = {1024, 1024, 1024, 1024}, tiles = {32, 32, 32};
  std::string mode = "strict";
  bool bf16 = true, bias = true, relu = true, kernel_const = true;
};

// the flags of an mlir-gen command line, with mlir-gen's defaults for what is absent
bool parse_mlir_gen(const std::string &line, Options &o) {
  o.bf16 = false; o.bias = false; o.relu = false; o.kernel_const = true; o.vnni = -1;
  o.tiles = {32, 32, 32};
  size_t i = 0;
  while (i < line.size()) {
    while (i < line.size() && isspace((unsigned char)line[i])) ++i;
    size_t j = i;
    while (j < line.size() && !isspace((unsigned char)line[j])) ++j;
    if (j == i) break;
    const std::string tok = line.substr(i, j - i);
    i = j;
    const size_t eq = tok.find('=');
    const std::string key = tok.substr(0, eq), val = eq == std::string::npos ? "" : tok.substr(eq + 1);
    if (key == "--bias") o.bias = true;
    else if (key == "--relu") o.relu = true;
    else if (key == "--kernel") o.kernel_const = val != "args";
    else if (key == "--float-type") {
      if (val != "bf16" && val != "f32") return false;
      o.bf16 = val == "bf16";
    } else if (key == "--vnni") o.vnni = atoll(val.c_str());
    else if (key == "--batch") o.batch = atoll(val.c_str());
    else if (key == "--layers") o.layers = ints(val.c_str());
    else if (key == "--tiles") o.tiles = ints(val.c_str());
    else if (key == "--seed") o.seed = atoll(val.c_str()) ? atoll(val.c_str()) : o.seed;
    else if (key == "--output" || key == "--keep-generic-matmul" || key == "mlir-gen") continue;   // no effect on the call stream
    else return false;
  }
  if (o.vnni < 0) o.vnni = o.bf16 ? 2 : 0;   // the pipeline packs bf16 weights with the factor the library reports
  if (!o.bf16) o.vnni = 0;
  return true;
}
}  // namespace

int main(int argc, char **argv) {
  Options o;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string a = argv[i];
    if (a == "--batch") o.batch = atoll(argv[i + 1]);
    else if (a == "--layers") o.layers = ints(argv[i + 1]);
    else if (a == "--tiles") o.tiles = ints(argv[i + 1]);
    else if (a == "--vnni") o.vnni = atoll(argv[i + 1]);
    else if (a == "-n") o.n_iter = atoll(argv[i + 1]);
    else if (a == "--seed") o.seed = atoll(argv[i + 1]);
    else if (a == "--mode") o.mode = argv[i + 1];
    else if (a == "--float-type") {
      const std::string v = argv[i + 1];
      if (v != "bf16" && v != "f32") { fprintf(stderr, "bad --float-type %s\n", v.c_str()); return 2; }
      o.bf16 = v == "bf16";
      if (!o.bf16) o.vnni = 0;
    }
    else if (a == "--bias") o.bias = atoll(argv[i + 1]) != 0;
    else if (a == "--relu") o.relu = atoll(argv[i + 1]) != 0;
    else if (a == "--kernel") o.kernel_const = std::string(argv[i + 1]) != "args";
    else if (a == "--mlir-gen") {
      if (!parse_mlir_gen(argv[i + 1], o)) { fprintf(stderr, "cannot read the mlir-gen flags: %s\n", argv[i + 1]); return 2; }
    }
    else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
  }
  const int64_t batch = o.batch, n_iter = o.n_iter, vnni = o.vnni, seed = o.seed;
  const std::vector<int64_t> &layers = o.layers, &tiles = o.tiles;
  const std::string &mode = o.mode;
  const bool bf16 = o.bf16, fused = o.bias || o.relu;
  const int64_t dtype = bf16 ? 2 : 1, es = bf16 ? 2 : 4;
  if (tiles.size() != 3 || layers.size() < 2 || (vnni != 0 && vnni != 2 && vnni != 4) || (vnni && !bf16)) {
    fprintf(stderr, "bad shape options\n");
    return 2;
  }
  const int64_t bn = tiles[0], bk = tiles[1], bc = tiles[2];
  const int64_t L = (int64_t)layers.size() - 1;
  if (batch % bn) { fprintf(stderr, "batch must be a multiple of the bn tile\n"); return 2; }
  for (int64_t l = 0; l < L; ++l)
    if (layers[l] % bc || layers[l + 1] % bk) { fprintf(stderr, "layer sizes must be multiples of the tiles\n"); return 2; }
  if (vnni && bc % vnni) { fprintf(stderr, "the bc tile must be a multiple of the VNNI factor\n"); return 2; }
  if (L > 1 && bk != bc) { fprintf(stderr, "multi-layer nets need bk == bc\n"); return 2; }
  // mlir-gen --vnni=4: the compiler asks the library for the packing factor (VNNIUtils.cpp:31-37); this is how it answers 4
  if (vnni == 4) setenv("TPP_XSMM_VNNI", "4", 1);

  // ---- arguments: splat constants first in op order (W1, b1, W2, b2, ...), then the kernel argument (MLIRBench.cpp:111-164)
  NormalInit init((int)seed);
  std::vector<void *> W(L), B(L), act(L + 1);
  for (int64_t l = 0; l < L; ++l) {
    W[l] = host_alloc(layers[l] * layers[l + 1], es);   // block-packed [K/bk][C/bc][bc][bk] (VNNI: [..][bc/v][bk][v]): same element count
    B[l] = host_alloc(layers[l + 1], es);
    init.fill(W[l], layers[l] * layers[l + 1], bf16);
    if (o.bias) init.fill(B[l], layers[l + 1], bf16);
  }
  for (int64_t l = 0; l <= L; ++l) act[l] = host_alloc(batch * layers[l], es);
  init.fill(act[0], batch * layers[0], bf16);

  // ---- dispatch, hoisted out of the loops (one handle: every layer has the same tile shape)
  // (bf16: the AMX tile-config pass adds the two no-reset / no-setup bits and wraps every invoke, IntelAMXTileConfig.cpp:32-139)
  const int64_t gflags = XSMM_GEMM_FLAG_BETA_0 | (vnni ? XSMM_GEMM_FLAG_ROWMAJOR_B_VNNI : 0) | (bf16 ? 64 | 128 : 0);
  const int64_t h = fused ? xsmm_fused_brgemm_dispatch(dtype, bn, bk, bc, bc, bk, bk, bn * bc, bc * bk, gflags, 0,
                                                        o.relu ? XSMM_UNARY_RELU : XSMM_UNARY_NONE,
                                                        o.bias ? XSMM_BINARY_FLAG_BCAST_COL_IN_0 : 0,
                                                        o.bias ? XSMM_BINARY_ADD : XSMM_BINARY_NONE)
                          : xsmm_brgemm_dispatch(dtype, bn, bk, bc, bc, bk, bk, bn * bc, bc * bk, gflags);
  const int64_t tc = bf16 ? xsmm_intel_amx_tile_config_dispatch(dtype, bn, bk, bc, bc, bk, bk, bn * bc, bc * bk, gflags) : 0;

  // ---- residency
  std::vector<void *> dW = W, dB = B, dact = act;
  const bool device = mode != "strict";
  if (device) {   // gpu.alloc + gpu.memcpy of every argument (MLIRBench::registerOnGpu)
    for (int64_t l = 0; l < L; ++l) {
      xsmm_cuda_register_host(W[l], layers[l] * layers[l + 1] * es, 1);
      xsmm_cuda_register_host(B[l], layers[l + 1] * es, 1);
      dW[l] = xsmm_cuda_device_ptr(W[l]);
      dB[l] = xsmm_cuda_device_ptr(B[l]);
    }
    for (int64_t l = 0; l <= L; ++l) {
      xsmm_cuda_register_host(act[l], batch * layers[l] * es, 1);
      dact[l] = xsmm_cuda_device_ptr(act[l]);
      // --kernel=const: the outputs of all layers but the last are buffers the generated kernel allocates and frees itself
      // (patches/0006 marks them); --kernel=args: every output is an argument of the kernel and stays an ordinary buffer
      if (o.kernel_const && l > 0 && l < L) xsmm_cuda_mark_temporary(dact[l], batch * layers[l] * es);
    }
  }

  if (mode == "lazy") xsmm_cuda_set_lazy(1);
  char amx_state[64];
  auto kernel = [&]() {   // the loop nest the pipeline emits (scf.parallel serialised on one stream)
    for (int64_t l = 0; l < L; ++l) {
      const int64_t nb_c = layers[l] / bc, nb_k = layers[l + 1] / bk;
      for (int64_t in = 0; in < batch / bn; ++in)
        for (int64_t ik = 0; ik < nb_k; ++ik) {
          if (bf16) xsmm_intel_amx_tile_config_invoke(dtype, tc, amx_state, 0);
          if (fused)
            xsmm_fused_brgemm_invoke(dtype, h, dact[l], in * nb_c * bn * bc, dW[l], ik * nb_c * bc * bk, dact[l + 1],
                                     (in * nb_k + ik) * bn * bk, o.bias ? dB[l] : nullptr, o.bias ? ik * bk : 0, nb_c);
          else
            xsmm_brgemm_invoke(dtype, h, dact[l], in * nb_c * bn * bc, dW[l], ik * nb_c * bc * bk, dact[l + 1],
                               (in * nb_k + ik) * bn * bk, nb_c);
          if (bf16) xsmm_intel_amx_tile_config_invoke(dtype, tc, amx_state, 0);
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
    xsmm_cuda_update_host(act[L], batch * layers[L] * es);
    xsmm_cuda_sync();
  }
  double checksum = 0.0;
  for (int64_t i = 0; i < batch * layers[L]; ++i)
    checksum += bf16 ? bf16_to_f32(static_cast<uint16_t *>(act[L])[i]) : static_cast<float *>(act[L])[i];
  double flops = 0.0;   // mlir-gen's BENCH_TOTAL_FLOPS (tools/mlir-gen/MLIRGen.cpp:313-334): 2MNK, + MN per bias add / relu
  for (int64_t l = 0; l < L; ++l)
    flops += 2.0 * batch * layers[l] * layers[l + 1] + ((o.bias ? 1.0 : 0.0) + (o.relu ? 1.0 : 0.0)) * batch * layers[l + 1];
  printf("{\"mode\": \"%s\", \"seconds_per_iteration\": %.9f, \"gflops\": %.3f, \"iterations\": %lld, \"kernel\": \"%s\", "
         "\"launches\": %lld, \"checksum\": %.6f, \"float_type\": \"%s\", \"vnni\": %lld, \"bias\": %d, \"relu\": %d, "
         "\"total_flops\": %.0f}\n",
         mode.c_str(), mean, flops / mean / 1e9, (long long)n_iter, xsmm_cuda_last_kernel(), (long long)xsmm_cuda_launch_count(),
         checksum, bf16 ? "bf16" : "f32", (long long)vnni, (int)o.bias, (int)o.relu, flops);
  if (graph) xsmm_cuda_graph_destroy(graph);
  return 0;
}
