// xsmm_abi_stub.cpp - TEST INFRASTRUCTURE: a recording stand-in for the part of the C-ABI that
// tpp_mlir_b200/csrc/harness/replay.cpp calls, so that the host-side replay logic (which graphs are captured, in which
// order invokes / copies / waits are issued) can be checked on a box without a GPU. It computes nothing.
#include <cstdint>
#include <vector>

#include "tpp_xsmm_abi.h"

namespace {
struct Event {
  int64_t kind;   // 1 invoke, 2 upload_async, 3 download_async, 4 wait_host, 5 stream_sync, 6 update_device, 7 update_host,
                  // 8 graph_launch marker (a = graph id)
  int64_t a, b, c, d, e;
};
std::vector<Event> g_log;                       // what "the device" was asked to do, in order
std::vector<std::vector<Event>> g_graphs;       // captured bodies, graph id = index + 1
std::vector<Event> g_capture;
bool g_capturing = false;
void *g_stream = nullptr;
int64_t g_streams_created = 0;
int64_t g_wait_ring = 1 << 30;                  // how many of the most recent downloads wait_host still knows
std::vector<int64_t> g_downloads;               // hosts of the downloads issued outside captures, in order

void record(const Event &e) { (g_capturing ? g_capture : g_log).push_back(e); }
}  // namespace

extern "C" {
void xsmm_fused_brgemm_invoke(int64_t dtype, int64_t addr, void *A, int64_t offA, void *B, int64_t offB, void *C,
                              int64_t offC, void *D, int64_t offD, int64_t numBatches) {
  (void)dtype;
  record({1, addr, (int64_t)(intptr_t)A + 2 * offA, (int64_t)(intptr_t)B + 2 * offB, (int64_t)(intptr_t)C + 2 * offC,
          D ? (int64_t)(intptr_t)D + 2 * offD : 0});
  (void)numBatches;
}
int64_t xsmm_cuda_upload_async(void *host, int64_t bytes) { record({2, (int64_t)(intptr_t)host, bytes, 0, 0, 0}); return 0; }
int64_t xsmm_cuda_download_async(void *host, int64_t bytes) {
  record({3, (int64_t)(intptr_t)host, bytes, 0, 0, 0});
  if (!g_capturing) g_downloads.push_back((int64_t)(intptr_t)host);
  return 0;
}
int64_t xsmm_cuda_wait_host(void *host) {
  record({4, (int64_t)(intptr_t)host, 0, 0, 0, 0});
  const int64_t n = (int64_t)g_downloads.size();
  for (int64_t i = n - 1; i >= 0 && i >= n - g_wait_ring; --i)
    if (g_downloads[(size_t)i] == (int64_t)(intptr_t)host) return 0;
  return -1;   // like the runtime: no such download on record any more
}
void xsmm_cuda_stream_sync(void) { record({5, (int64_t)(intptr_t)g_stream, 0, 0, 0, 0}); }
int64_t xsmm_cuda_update_device(void *host, int64_t bytes) { record({6, (int64_t)(intptr_t)host, bytes, 0, 0, 0}); return 0; }
int64_t xsmm_cuda_update_host(void *host, int64_t bytes) { record({7, (int64_t)(intptr_t)host, bytes, 0, 0, 0}); return 0; }
int64_t xsmm_cuda_graph_begin(void) {
  if (g_capturing) return -1;
  g_capturing = true;
  g_capture.clear();
  return 0;
}
int64_t xsmm_cuda_graph_end(void) {
  if (!g_capturing) return 0;
  g_capturing = false;
  g_graphs.push_back(g_capture);
  return (int64_t)g_graphs.size();
}
void xsmm_cuda_graph_launch(int64_t graph) {
  g_log.push_back({8, graph, 0, 0, 0, 0});
  const std::vector<Event> &body = g_graphs[(size_t)graph - 1];
  g_log.insert(g_log.end(), body.begin(), body.end());
}
void *xsmm_cuda_get_stream(void) { return g_stream; }
void xsmm_cuda_set_stream(void *s) { g_stream = s; }
void *xsmm_cuda_stream_create(void) { return (void *)(intptr_t)(0x1000 + ++g_streams_created); }

// ---- test access ----
__attribute__((visibility("default"))) void stub_reset(void) {
  g_log.clear();
  g_graphs.clear();
  g_capture.clear();
  g_capturing = false;
  g_downloads.clear();
  g_wait_ring = 1 << 30;
}
__attribute__((visibility("default"))) void stub_set_wait_ring(int64_t n) { g_wait_ring = n; }
__attribute__((visibility("default"))) int64_t stub_log_size(void) { return (int64_t)g_log.size(); }
__attribute__((visibility("default"))) void stub_log_get(int64_t i, int64_t *out6) {
  const Event &e = g_log[(size_t)i];
  out6[0] = e.kind; out6[1] = e.a; out6[2] = e.b; out6[3] = e.c; out6[4] = e.d; out6[5] = e.e;
}
__attribute__((visibility("default"))) int64_t stub_num_graphs(void) { return (int64_t)g_graphs.size(); }
}
