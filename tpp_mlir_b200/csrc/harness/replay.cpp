// replay.cpp - native stand-in for the loop tpp-run's JIT-compiled main() executes.
//
// tpp-run lowers the benchmark kernel to LLVM IR whose hot loop is a plain native loop of
// func.call @xsmm_fused_brgemm_invoke(...) (lib/TPP/Runner/MLIRBench.cpp
// :265-295 wraps it in perf.bench). LLVM/MLIR are not available here, so this file is that loop
// written by hand: it calls ONLY the C-ABI of include/tpp_xsmm_abi.h, exactly as JIT'd code
// would, for the block-packed MLP that `mlir-gen --kernel=const --bias --relu` generates
// (tools/mlir-gen/MLIRGen.cpp:632-681; offsets per SURVEY.md Appendix B).
#include <algorithm>
#include <cstdint>
#include <map>
#include <mutex>
#include <tuple>

#include "tpp_xsmm_abi.h"

namespace {
// the slot's previous output has reached the host; a download the runtime no longer remembers (its ring holds the
// most recent ones) is waited for by draining the thread's streams - never by reusing the buffer blindly
inline void wait_output(void *host) {
  if (xsmm_cuda_wait_host(host) != 0) xsmm_cuda_stream_sync();
}
}  // namespace

extern "C" {

struct TppMlpSet {       // one set of buffers (several sets are rotated to defeat the L2)
  void *acts[9];         // acts[0] = input, acts[l+1] = output of layer l
  void *weights[8];
  void *biases[8];
};

// Runs `steps` forward passes; step s uses sets[s % num_sets]. Layer l: for every (iN, iK)
// output block one fused_brgemm invoke with numBatches = C/bc (the scf.parallel loop nest of
// the reference, serialised on one CUDA stream).
__attribute__((visibility("default")))
void tpp_replay_mlp(int64_t dtype, int64_t num_layers, const int64_t *handles, const int64_t *layer_sizes,
                    int64_t batch, int64_t bn, int64_t bk, int64_t bc, const TppMlpSet *sets, int64_t num_sets,
                    int64_t first_step, int64_t steps, int64_t has_bias) {
  for (int64_t s = 0; s < steps; ++s) {
    const TppMlpSet &set = sets[(first_step + s) % num_sets];
    for (int64_t l = 0; l < num_layers; ++l) {
      const int64_t c = layer_sizes[l], k = layer_sizes[l + 1];
      const int64_t nb_c = c / bc, nb_k = k / bk;
      for (int64_t in = 0; in < batch / bn; ++in)
        for (int64_t ik = 0; ik < nb_k; ++ik)
          xsmm_fused_brgemm_invoke(dtype, handles[l], set.acts[l], in * nb_c * bn * bc, set.weights[l],
                                   ik * nb_c * bc * bk, set.acts[l + 1], (in * nb_k + ik) * bn * bk,
                                   has_bias ? set.biases[l] : nullptr, ik * bk, nb_c);
    }
  }
}

// Same loop, but the invoke sequence is captured once into CUDA graphs (xsmm_cuda_graph_begin/end around
// the invokes) and replayed: what a perf.bench lowering that captures its body would execute.
// graphs[i] (i < num_sets) replays one forward pass on operand set i; graphs[num_sets] replays one full
// rotation (num_sets consecutive forward passes, used when `group` != 0 - the loop body unrolled over the
// rotating operand sets, so the per-graph-launch latency is paid once per rotation). 0 == not captured yet.
// With `group`, the steps before the first rotation boundary and after the last one are replayed as ONE graph each
// (captured per (first set, length) on first use), single steps only when one step is left. `graphs` has
// num_sets + 2 entries: [num_sets + 1] identifies the caller's loop object for those partial graphs.
__attribute__((visibility("default")))
int64_t tpp_replay_mlp_graph(int64_t dtype, int64_t num_layers, const int64_t *handles, const int64_t *layer_sizes,
                             int64_t batch, int64_t bn, int64_t bk, int64_t bc, const TppMlpSet *sets,
                             int64_t num_sets, int64_t *graphs, int64_t first_step, int64_t steps, int64_t has_bias,
                             int64_t group) {
  int64_t s = 0;
  while (s < steps) {
    const int64_t idx = (first_step + s) % num_sets;
    if (group && idx == 0 && steps - s >= num_sets) {
      if (!graphs[num_sets]) {
        if (xsmm_cuda_graph_begin() != 0) return -1;
        tpp_replay_mlp(dtype, num_layers, handles, layer_sizes, batch, bn, bk, bc, sets, num_sets, 0, num_sets,
                       has_bias);
        graphs[num_sets] = xsmm_cuda_graph_end();
        if (!graphs[num_sets]) return -1;
      }
      xsmm_cuda_graph_launch(graphs[num_sets]);
      s += num_sets;
      continue;
    }
    // a partial rotation (the head up to the next rotation boundary, or the tail of the run): one captured graph per
    // (first set, length), so that a run of any length is still a handful of graph launches
    const int64_t chunk = std::min(steps - s, num_sets - idx);
    if (group && chunk >= 2) {
      // keyed by an id that lives in the caller's graphs array (slot num_sets + 1, 0 = not assigned yet), not by
      // addresses: a later loop object whose arrays land on recycled addresses must not inherit these graphs
      static std::map<std::tuple<int64_t, int64_t, int64_t>, int64_t> partial;
      static int64_t next_loop_id = 0;
      static std::mutex partial_mutex;
      int64_t g = 0;
      {
        std::lock_guard<std::mutex> lock(partial_mutex);
        int64_t &loop_id = graphs[num_sets + 1];
        if (!loop_id) loop_id = ++next_loop_id;
        g = partial[std::make_tuple(loop_id, idx, chunk)];
      }
      if (!g) {
        if (xsmm_cuda_graph_begin() != 0) return -1;
        tpp_replay_mlp(dtype, num_layers, handles, layer_sizes, batch, bn, bk, bc, sets + idx, chunk, 0, chunk, has_bias);
        g = xsmm_cuda_graph_end();
        if (!g) return -1;
        std::lock_guard<std::mutex> lock(partial_mutex);
        partial[std::make_tuple(graphs[num_sets + 1], idx, chunk)] = g;
      }
      xsmm_cuda_graph_launch(g);
      s += chunk;
      continue;
    }
    if (!graphs[idx]) {
      if (xsmm_cuda_graph_begin() != 0) return -1;
      tpp_replay_mlp(dtype, num_layers, handles, layer_sizes, batch, bn, bk, bc, sets + idx, 1, 0, 1, has_bias);
      graphs[idx] = xsmm_cuda_graph_end();
      if (!graphs[idx]) return -1;
    }
    xsmm_cuda_graph_launch(graphs[idx]);
    ++s;
  }
  return 0;
}

// The benchmark loop of tpp-run itself: `steps` forward passes on ONE set of buffers, back to back. graphs[0] replays
// `unroll` consecutive forward passes captured as one graph (the loop body unrolled before capture: the runtime runs the
// exact repeats as one launch, a plain sequence of layer passes), graphs[1] one forward pass (the remainder).
__attribute__((visibility("default")))
int64_t tpp_replay_mlp_graph_unrolled(int64_t dtype, int64_t num_layers, const int64_t *handles, const int64_t *layer_sizes,
                                      int64_t batch, int64_t bn, int64_t bk, int64_t bc, const TppMlpSet *set,
                                      int64_t *graphs, int64_t unroll, int64_t steps, int64_t has_bias) {
  if (unroll < 1) unroll = 1;
  int64_t s = 0;
  if (unroll > 1 && steps >= unroll) {
    if (!graphs[0]) {
      if (xsmm_cuda_graph_begin() != 0) return -1;
      tpp_replay_mlp(dtype, num_layers, handles, layer_sizes, batch, bn, bk, bc, set, 1, 0, unroll, has_bias);
      if (!(graphs[0] = xsmm_cuda_graph_end())) return -1;
    }
    for (; s + unroll <= steps; s += unroll) xsmm_cuda_graph_launch(graphs[0]);
  }
  if (s < steps) {
    if (!graphs[1]) {
      if (xsmm_cuda_graph_begin() != 0) return -1;
      tpp_replay_mlp(dtype, num_layers, handles, layer_sizes, batch, bn, bk, bc, set, 1, 0, 1, has_bias);
      if (!(graphs[1] = xsmm_cuda_graph_end())) return -1;
    }
    for (; s < steps; ++s) xsmm_cuda_graph_launch(graphs[1]);
  }
  return 0;
}

// One forward on host buffers that were registered with xsmm_cuda_register_host: upload the
// step's input, run the layers on the mirrors, download the step's output, wait for it.
// use_graph != 0: the whole step (H2D copy, invokes, D2H copy) is captured once with
// xsmm_cuda_graph_begin/end and replayed - one host call + one stream wait per step.
__attribute__((visibility("default")))
int64_t tpp_replay_mlp_e2e(int64_t dtype, int64_t num_layers, const int64_t *handles, const int64_t *layer_sizes,
                           int64_t batch, int64_t bn, int64_t bk, int64_t bc, const TppMlpSet *set, int64_t steps,
                           int64_t has_bias, int64_t elem_size, int64_t use_graph, int64_t *graph_io) {
  auto one_step = [&]() {
    xsmm_cuda_update_device(set->acts[0], batch * layer_sizes[0] * elem_size);
    tpp_replay_mlp(dtype, num_layers, handles, layer_sizes, batch, bn, bk, bc, set, 1, 0, 1, has_bias);
    xsmm_cuda_update_host(set->acts[num_layers], batch * layer_sizes[num_layers] * elem_size);
  };
  if (use_graph && graph_io && !*graph_io) {
    if (xsmm_cuda_graph_begin() != 0) return -1;
    one_step();
    *graph_io = xsmm_cuda_graph_end();
    if (!*graph_io) return -1;
  }
  for (int64_t s = 0; s < steps; ++s) {
    if (use_graph) xsmm_cuda_graph_launch(*graph_io);
    else one_step();
    xsmm_cuda_stream_sync();   // the step's output is in host memory
  }
  return 0;
}

// Throughput forms of the same end-to-end step: `depth` independent steps in flight, one set of host buffers +
// mirrors per slot; step s reuses slot s % depth only after that slot's previous output has reached host memory
// (the point where a caller would consume the result and write the next input). Every step still uploads its input
// and downloads its output inside the loop.
//   mode 0  one compute stream + the library's copy streams: upload_async(in) -> replay of the captured layer
//           sequence -> download_async(out) per step, wait_host(out) before the slot is reused. Kernels of
//           different steps never overlap (one stream); copies run under them.
//   mode 1  the loop body unrolled `depth` times inside ONE captured graph (the copies become parallel branches):
//           one graph launch + one stream wait per `depth` steps.
//   mode 2  one stream (xsmm_cuda_stream_create) and one captured step graph (copies included) per slot.
//   mode 2 + g (g >= 1)  groups of g steps: upload_async of the g inputs -> ONE captured graph with the g layer
//           sequences (independent chains, which the runtime fuses into one interleaved launch) -> download_async of the
//           g outputs; depth / g groups in flight, wait_host before a slot is reused.
// graphs: depth + 1 entries, streams: depth entries, 0 = not created yet; they persist across calls (one mode per
// pair of arrays). steps is rounded down to a multiple of depth in mode 1. Returns the number of steps run or -1.
__attribute__((visibility("default")))
int64_t tpp_replay_mlp_e2e_pipelined(int64_t dtype, int64_t num_layers, const int64_t *handles,
                                     const int64_t *layer_sizes, int64_t batch, int64_t bn, int64_t bk, int64_t bc,
                                     const TppMlpSet *slots, int64_t depth, int64_t steps, int64_t has_bias,
                                     int64_t elem_size, int64_t mode, int64_t *graphs, void **streams) {
  const int64_t in_bytes = batch * layer_sizes[0] * elem_size;
  const int64_t out_bytes = batch * layer_sizes[num_layers] * elem_size;
  auto layers = [&](int64_t d) {
    tpp_replay_mlp(dtype, num_layers, handles, layer_sizes, batch, bn, bk, bc, slots + d, 1, 0, 1, has_bias);
  };
  if (mode == 0) {
    for (int64_t d = 0; d < depth; ++d) {
      if (graphs[d]) continue;
      if (xsmm_cuda_graph_begin() != 0) return -1;
      layers(d);
      if (!(graphs[d] = xsmm_cuda_graph_end())) return -1;
    }
    for (int64_t s = 0; s < steps; ++s) {
      const int64_t d = s % depth;
      void *out = slots[d].acts[num_layers];
      if (s >= depth) wait_output(out);   // output of step s - depth consumed; the slot is free again
      xsmm_cuda_upload_async(slots[d].acts[0], in_bytes);
      xsmm_cuda_graph_launch(graphs[d]);
      xsmm_cuda_download_async(out, out_bytes);
    }
    xsmm_cuda_stream_sync();   // drain: every step's output has reached the host
    return steps;
  }
  if (mode == 1) {
    if (!graphs[depth]) {
      if (xsmm_cuda_graph_begin() != 0) return -1;
      for (int64_t d = 0; d < depth; ++d) {
        xsmm_cuda_upload_async(slots[d].acts[0], in_bytes);
        layers(d);
        xsmm_cuda_download_async(slots[d].acts[num_layers], out_bytes);
      }
      if (!(graphs[depth] = xsmm_cuda_graph_end())) return -1;
    }
    const int64_t groups = steps / depth;
    for (int64_t g = 0; g < groups; ++g) {
      xsmm_cuda_graph_launch(graphs[depth]);
      xsmm_cuda_stream_sync();   // the outputs of these `depth` steps are in host memory
    }
    return groups * depth;
  }
  if (mode >= 3) {
    // mode 3 + g: like mode 0, but the unit of work is a GROUP of g = mode - 2 ... consecutive steps (slots): their
    // inputs are uploaded, their layer sequences replay as ONE captured graph (independent chains: the runtime runs
    // them interleaved in one launch), their outputs are downloaded; depth / g groups are in flight.
    const int64_t gsz = mode - 2;                 // mode 3 -> 1 (== mode 0), mode 5 -> 3 steps per launch, ...
    const int64_t ngrp = depth / gsz;
    if (ngrp < 1) return -1;
    for (int64_t grp = 0; grp < ngrp; ++grp) {
      if (graphs[grp]) continue;
      if (xsmm_cuda_graph_begin() != 0) return -1;
      for (int64_t j = 0; j < gsz; ++j) layers(grp * gsz + j);
      if (!(graphs[grp] = xsmm_cuda_graph_end())) return -1;
    }
    // a group whose input (output) buffers are consecutive slices of ONE registered host block moves them with one
    // copy: a 512 KiB copy reaches ~39 GB/s on this PCIe link, a 37 MiB copy ~55 GB/s (scripts/pcie_probe.py)
    auto contiguous = [&](int64_t grp, int64_t which, int64_t bytes) {
      for (int64_t j = 1; j < gsz; ++j)
        if (static_cast<char *>(slots[grp * gsz + j].acts[which]) !=
            static_cast<char *>(slots[grp * gsz].acts[which]) + j * bytes)
          return false;
      return true;
    };
    const int64_t iters = steps / gsz;
    for (int64_t it = 0; it < iters; ++it) {
      const int64_t grp = it % ngrp;
      const TppMlpSet &g0 = slots[grp * gsz];
      const bool in_block = contiguous(grp, 0, in_bytes), out_block = contiguous(grp, num_layers, out_bytes);
      // previous outputs of this group's slots consumed: the buffers are free again
      if (it >= ngrp) {
        if (out_block) wait_output(g0.acts[num_layers]);
        else for (int64_t j = 0; j < gsz; ++j) wait_output(slots[grp * gsz + j].acts[num_layers]);
      }
      if (in_block) xsmm_cuda_upload_async(g0.acts[0], gsz * in_bytes);
      else for (int64_t j = 0; j < gsz; ++j) xsmm_cuda_upload_async(slots[grp * gsz + j].acts[0], in_bytes);
      xsmm_cuda_graph_launch(graphs[grp]);
      if (out_block) xsmm_cuda_download_async(g0.acts[num_layers], gsz * out_bytes);
      else for (int64_t j = 0; j < gsz; ++j) xsmm_cuda_download_async(slots[grp * gsz + j].acts[num_layers], out_bytes);
    }
    xsmm_cuda_stream_sync();
    return iters * gsz;
  }
  void *caller_stream = xsmm_cuda_get_stream();
  for (int64_t d = 0; d < depth; ++d) {
    if (!streams[d]) streams[d] = xsmm_cuda_stream_create();
    if (!graphs[d]) {
      xsmm_cuda_set_stream(streams[d]);
      if (xsmm_cuda_graph_begin() != 0) return -1;
      xsmm_cuda_update_device(slots[d].acts[0], in_bytes);
      layers(d);
      xsmm_cuda_update_host(slots[d].acts[num_layers], out_bytes);
      if (!(graphs[d] = xsmm_cuda_graph_end())) return -1;
    }
  }
  for (int64_t s = 0; s < steps; ++s) {
    const int64_t d = s % depth;
    xsmm_cuda_set_stream(streams[d]);
    if (s >= depth) xsmm_cuda_stream_sync();
    xsmm_cuda_graph_launch(graphs[d]);
  }
  for (int64_t d = 0; d < depth; ++d) {
    xsmm_cuda_set_stream(streams[d]);
    xsmm_cuda_stream_sync();
  }
  xsmm_cuda_set_stream(caller_stream);
  return steps;
}

} // extern "C"
