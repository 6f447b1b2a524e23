// xsmm_abi_stub_standin.cpp - TEST INFRASTRUCTURE: a recording stand-in for the part of the C-ABI that
// tpp_mlir_b200/csrc/harness/tpp_run_standin.cpp calls, so that WHAT the stand-in asks of the runtime for a given
// mlir-gen command line (dispatch arguments, invoke counts and offsets, which buffers are registered / marked temporary,
// graph use) can be checked on a box without a GPU. It computes nothing. At exit it writes one JSON object to $STUB_LOG.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "tpp_xsmm_abi.h"

namespace {
struct Log {
  std::vector<std::vector<int64_t>> dispatches;   // [kind (0 brgemm, 1 fused, 2 tile config), args...]
  std::vector<std::vector<int64_t>> first_invokes;   // the first 4 compute invokes, [kind, dtype, offA, offB, offC, offD, hasD, batch]
  int64_t brgemm_invokes = 0, fused_invokes = 0, tilecfg_invokes = 0;
  int64_t registered = 0, registered_bytes = 0, temporaries = 0, temporary_bytes = 0;
  int64_t graphs = 0, graph_launches = 0, captured_invokes = 0, lazy = 0, timers = 0;
  int64_t max_off_a = 0, max_off_b = 0, max_off_c = 0, max_off_d = 0;
  bool capturing = false;
  std::string vnni_env_at_dispatch;
  ~Log() {
    const char *path = getenv("STUB_LOG");
    FILE *f = path ? fopen(path, "w") : nullptr;
    if (!f) return;
    auto arr = [&](const std::vector<std::vector<int64_t>> &v) {
      fputc('[', f);
      for (size_t i = 0; i < v.size(); ++i) {
        fputs(i ? ", [" : "[", f);
        for (size_t j = 0; j < v[i].size(); ++j) fprintf(f, "%s%lld", j ? ", " : "", (long long)v[i][j]);
        fputc(']', f);
      }
      fputc(']', f);
    };
    fputs("{\"dispatches\": ", f); arr(dispatches);
    fputs(", \"first_invokes\": ", f); arr(first_invokes);
    fprintf(f, ", \"brgemm_invokes\": %lld, \"fused_invokes\": %lld, \"tilecfg_invokes\": %lld, \"registered\": %lld, "
               "\"registered_bytes\": %lld, \"temporaries\": %lld, \"temporary_bytes\": %lld, \"graphs\": %lld, "
               "\"graph_launches\": %lld, \"captured_invokes\": %lld, \"lazy\": %lld, \"timers\": %lld, "
               "\"max_off\": [%lld, %lld, %lld, %lld], \"vnni_env_at_dispatch\": \"%s\"}\n",
            (long long)brgemm_invokes, (long long)fused_invokes, (long long)tilecfg_invokes, (long long)registered,
            (long long)registered_bytes, (long long)temporaries, (long long)temporary_bytes, (long long)graphs,
            (long long)graph_launches, (long long)captured_invokes, (long long)lazy, (long long)timers, (long long)max_off_a,
            (long long)max_off_b, (long long)max_off_c, (long long)max_off_d, vnni_env_at_dispatch.c_str());
    fclose(f);
  }
} g;

void note_invoke(int64_t kind, int64_t dtype, int64_t oa, int64_t ob, int64_t oc, int64_t od, bool has_d, int64_t batch) {
  if (g.capturing) ++g.captured_invokes;
  if (g.first_invokes.size() < 4) g.first_invokes.push_back({kind, dtype, oa, ob, oc, od, has_d ? 1 : 0, batch});
  if (oa > g.max_off_a) g.max_off_a = oa;
  if (ob > g.max_off_b) g.max_off_b = ob;
  if (oc > g.max_off_c) g.max_off_c = oc;
  if (od > g.max_off_d) g.max_off_d = od;
}
void note_env() {
  const char *e = getenv("TPP_XSMM_VNNI");
  g.vnni_env_at_dispatch = e ? e : "";
}
}  // namespace

extern "C" {
int64_t xsmm_brgemm_dispatch(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc,
                             int64_t sa, int64_t sb, int64_t flags) {
  note_env();
  g.dispatches.push_back({0, dtype, m, n, k, lda, ldb, ldc, sa, sb, flags});
  return 0x1000 + (int64_t)g.dispatches.size();
}
int64_t xsmm_fused_brgemm_dispatch(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc,
                                   int64_t sa, int64_t sb, int64_t gf, int64_t uf, int64_t uk, int64_t bf, int64_t bk) {
  note_env();
  g.dispatches.push_back({1, dtype, m, n, k, lda, ldb, ldc, sa, sb, gf, uf, uk, bf, bk});
  return 0x1000 + (int64_t)g.dispatches.size();
}
int64_t xsmm_intel_amx_tile_config_dispatch(int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb,
                                            int64_t ldc, int64_t sa, int64_t sb, int64_t flags) {
  g.dispatches.push_back({2, dtype, m, n, k, lda, ldb, ldc, sa, sb, flags});
  return 0x1000 + (int64_t)g.dispatches.size();
}
void xsmm_brgemm_invoke(int64_t dtype, int64_t, void *, int64_t oa, void *, int64_t ob, void *, int64_t oc, int64_t batch) {
  ++g.brgemm_invokes;
  note_invoke(0, dtype, oa, ob, oc, 0, false, batch);
}
void xsmm_fused_brgemm_invoke(int64_t dtype, int64_t, void *, int64_t oa, void *, int64_t ob, void *, int64_t oc, void *D,
                              int64_t od, int64_t batch) {
  ++g.fused_invokes;
  note_invoke(1, dtype, oa, ob, oc, od, D != nullptr, batch);
}
void xsmm_intel_amx_tile_config_invoke(int64_t, int64_t, void *, int64_t) { ++g.tilecfg_invokes; }
int64_t perf_start_timer(void) { ++g.timers; return 1; }
double perf_stop_timer(int64_t) { return 1.0; }
void xsmm_cuda_sync(void) {}
int64_t xsmm_cuda_register_host(void *, int64_t bytes, int64_t) { ++g.registered; g.registered_bytes += bytes; return 0; }
int64_t xsmm_cuda_update_host(void *, int64_t) { return 0; }
void *xsmm_cuda_device_ptr(void *host) { return host; }
int64_t xsmm_cuda_graph_begin(void) { g.capturing = true; return 0; }
int64_t xsmm_cuda_graph_end(void) { g.capturing = false; return ++g.graphs; }
void xsmm_cuda_graph_launch(int64_t) { ++g.graph_launches; }
void xsmm_cuda_graph_destroy(int64_t) {}
void xsmm_cuda_set_lazy(int64_t on) { g.lazy = on; }
void xsmm_cuda_mark_temporary(void *, int64_t bytes) { ++g.temporaries; g.temporary_bytes += bytes; }
int64_t xsmm_cuda_launch_count(void) { return 0; }
const char *xsmm_cuda_last_kernel(void) { return "stub"; }
}
