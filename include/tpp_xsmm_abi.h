/*
 * tpp_xsmm_abi.h - C-ABI of the B200-native TPP execution backend.
 *
 * This is the drop-in boundary: the shared library built from
 * tpp_mlir_b200/csrc (libtpp_xsmm_runner_utils.so) exports exactly the symbols
 * that tpp-mlir's JIT-compiled code calls after the ConvertXsmmToFunc lowering.
 *
 * Reference interfaces replaced (all paths relative to the tpp-mlir tree):
 *   runtime/Xsmm/XsmmRunnerUtils.h:22-83     the 13 xsmm_* entry points
 *   runtime/PerfRunnerUtils.h:22-24          perf_start_timer / perf_stop_timer
 *   lib/TPP/Transforms/Utils/VNNIUtils.cpp:36 libxsmm_cpuid_dot_pack_factor
 *   lib/TPP/Conversion/ConvertXsmmToFunc/ConvertXsmmToFunc.cpp:37-101,298-352
 *                                            argument order / types (all i64)
 *
 * Differences from the reference header, on purpose:
 *   - every integer (including the enums) is declared int64_t, because that is
 *     what the lowering really passes (MLIR i64); the reference declares 32-bit
 *     enums and relies on the x86-64 SysV register width.
 *   - pointers may be DEVICE pointers (fast path, no copies), host pointers that
 *     were registered with xsmm_cuda_register_host (translated to a device
 *     mirror), or plain host pointers (strict mode: operands are staged to the
 *     GPU, the kernel runs, the result is copied back and the call returns
 *     after the result is visible to the host - the reference's semantics).
 *   - there is NO CPU fallback: if no CUDA device / sm_100 kernel image is
 *     available the first dispatch prints a diagnostic and exit(-1)s, the same
 *     way the reference does when libxsmm cannot JIT a kernel
 *     (runtime/Xsmm/XsmmRunnerUtils.cpp:132-137).
 *
 * Layout notation (row-major view, what tpp-mlir passes):
 *   A[b][i][p] at A + b*stride_a + i*lda + p      (i<m, p<k)
 *   B[b][p][j] at B + b*stride_b + p*ldb + j      (j<n)   or VNNI [K/2][N][2]
 *   C[i][j]    at C + i*ldc + j
 * All ld and stride values are in ELEMENTS. ptr = alignedPtr + offset*sizeof(T)
 * (runtime/Xsmm/XsmmRunnerUtils.cpp:63-75).
 */
#ifndef TPP_XSMM_ABI_H
#define TPP_XSMM_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TPP_XSMM_EXPORT __attribute__((visibility("default")))
#else
#define TPP_XSMM_EXPORT
#endif

/* ---- enum values (include/TPP/Dialect/Xsmm/XsmmEnum.td:13-84) ------------- */
enum {
  XSMM_DTYPE_F32 = 1,
  XSMM_DTYPE_BF16 = 2
};
enum {
  XSMM_BINARY_NONE = 0,
  XSMM_BINARY_ADD = 1,
  XSMM_BINARY_MUL = 2,
  XSMM_BINARY_SUB = 3,
  XSMM_BINARY_DIV = 4
};
enum {
  XSMM_UNARY_NONE = 0,
  XSMM_UNARY_IDENTITY = 1,
  XSMM_UNARY_ZERO = 2,
  XSMM_UNARY_RELU = 5,
  XSMM_UNARY_VNNI2 = 28,
  XSMM_UNARY_TRANSPOSE = 29,
  /* VNNI-4 pack [K][N] -> [K/4][N][4]: the libxsmm transform the reference runtime accepts (NORM_TO_VNNI4 in the
   * list of runtime/Xsmm/XsmmRunnerUtils.cpp:45-54; the dialect has no op for it yet, mlir-gen --vnni=4 packs weights at
   * compile time). 32 follows libxsmm's numbering after NORM_TO_VNNI2 = 28 / NORM_TO_NORMT = 29. */
  XSMM_UNARY_VNNI4 = 32,
  /* extensions (not in the reference dialect): inverses, [K/v][N][v] -> [K][N] */
  XSMM_UNARY_UNVNNI2_EXT = 1028,
  XSMM_UNARY_UNVNNI4_EXT = 1032
};
enum {
  XSMM_UNARY_FLAG_NONE = 0,
  XSMM_UNARY_FLAG_BCAST_ROW = 2,
  XSMM_UNARY_FLAG_BCAST_COL = 4,
  XSMM_UNARY_FLAG_BCAST_SCALAR = 8
};
enum {
  XSMM_BINARY_FLAG_NONE = 0,
  XSMM_BINARY_FLAG_BCAST_ROW_IN_0 = 1,
  XSMM_BINARY_FLAG_BCAST_ROW_IN_1 = 2,
  XSMM_BINARY_FLAG_BCAST_COL_IN_0 = 4,
  XSMM_BINARY_FLAG_BCAST_COL_IN_1 = 8,
  XSMM_BINARY_FLAG_BCAST_SCALAR_IN_0 = 16,
  XSMM_BINARY_FLAG_BCAST_SCALAR_IN_1 = 32
};
/* GEMM flags AS RECEIVED by the C-ABI, i.e. after the lowering swapped the
 * dialect's vnni_a <-> vnni_b for libxsmm's column-major view
 * (ConvertXsmmToFunc.cpp:251-265; test/Conversion/XsmmToFunc/xsmm-to-func.mlir:86-116):
 *   2048 (libxsmm VNNI_A) == dialect vnni_b : row-major B operand is [K/2][N][2]
 *   4096 (libxsmm VNNI_B) == dialect vnni_a : row-major A operand is [M][K/2][2]
 *                                             (bit-identical to the flat [M][K]) */
enum {
  XSMM_GEMM_FLAG_NONE = 0,
  XSMM_GEMM_FLAG_BETA_0 = 4,
  XSMM_GEMM_FLAG_NO_RESET_TILECONFIG = 64,
  XSMM_GEMM_FLAG_NO_SETUP_TILECONFIG = 128,
  XSMM_GEMM_FLAG_ROWMAJOR_B_VNNI = 2048,
  XSMM_GEMM_FLAG_ROWMAJOR_A_VNNI = 4096,
  XSMM_GEMM_FLAG_VNNI_C = 8192
};

/* ---- dispatch: build (or look up) a kernel descriptor, return it as i64 ---- */

/* replaces runtime/Xsmm/XsmmRunnerUtils.h:22-24 (XsmmRunnerUtils.cpp:95-140) */
TPP_XSMM_EXPORT int64_t xsmm_gemm_dispatch(int64_t dtype, int64_t m, int64_t n,
                                           int64_t k, int64_t lda, int64_t ldb,
                                           int64_t ldc, int64_t flags);

/* replaces XsmmRunnerUtils.h:26-28 (XsmmRunnerUtils.cpp:142-179) */
TPP_XSMM_EXPORT int64_t xsmm_unary_dispatch(int64_t kind, int64_t dtype,
                                            int64_t m, int64_t n, int64_t ldi,
                                            int64_t ldo, int64_t flags);

/* replaces XsmmRunnerUtils.h:30-32 (XsmmRunnerUtils.cpp:181-211) */
TPP_XSMM_EXPORT int64_t xsmm_binary_dispatch(int64_t kind, int64_t dtype,
                                             int64_t m, int64_t n,
                                             int64_t ldiLhs, int64_t ldiRhs,
                                             int64_t ldo, int64_t flags);

/* replaces XsmmRunnerUtils.h:34-36 (XsmmRunnerUtils.cpp:308-361) */
TPP_XSMM_EXPORT int64_t xsmm_brgemm_dispatch(int64_t dtype, int64_t m,
                                             int64_t n, int64_t k, int64_t lda,
                                             int64_t ldb, int64_t ldc,
                                             int64_t stride_a, int64_t stride_b,
                                             int64_t flags);

/* replaces XsmmRunnerUtils.h:38-45 (XsmmRunnerUtils.cpp:385-457) */
TPP_XSMM_EXPORT int64_t xsmm_fused_brgemm_dispatch(
    int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb,
    int64_t ldc, int64_t stride_a, int64_t stride_b, int64_t gemm_flags,
    int64_t unary_flags, int64_t unary_kind, int64_t binary_flags,
    int64_t binary_kind);

/* replaces XsmmRunnerUtils.h:47-49 (XsmmRunnerUtils.cpp:213-246): AMX tile
 * configuration has no meaning on a GPU; returns a non-zero dummy handle. */
TPP_XSMM_EXPORT int64_t xsmm_intel_amx_tile_config_dispatch(
    int64_t dtype, int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb,
    int64_t ldc, int64_t stride_a, int64_t stride_b, int64_t flags);

/* ---- invoke ---------------------------------------------------------------- */

/* replaces XsmmRunnerUtils.h:51-54 (XsmmRunnerUtils.cpp:79-93) */
TPP_XSMM_EXPORT void xsmm_gemm_invoke(int64_t dtype, int64_t addr,
                                      void *alignedPtrA, int64_t offsetA,
                                      void *alignedPtrB, int64_t offsetB,
                                      void *alignedPtrC, int64_t offsetC);

/* replaces XsmmRunnerUtils.h:56-59 (XsmmRunnerUtils.cpp:248-259) */
TPP_XSMM_EXPORT void xsmm_unary_invoke(int64_t dtype, int64_t addr,
                                       void *alignedPtrIn, int64_t offsetIn,
                                       void *alignedPtrOut, int64_t offsetOut);

/* replaces XsmmRunnerUtils.h:61-63 (XsmmRunnerUtils.cpp:276-286) */
TPP_XSMM_EXPORT void xsmm_unary_scalar_invoke(int64_t dtype, int64_t addr,
                                              float scalar, void *alignedPtrOut,
                                              int64_t offsetOut);

/* replaces XsmmRunnerUtils.h:65-68 (XsmmRunnerUtils.cpp:261-274) */
TPP_XSMM_EXPORT void xsmm_binary_invoke(int64_t dtype, int64_t addr,
                                        void *alignedPtrLhs, int64_t offsetLhs,
                                        void *alignedPtrRhs, int64_t offsetRhs,
                                        void *alignedPtrOut, int64_t offsetOut);

/* replaces XsmmRunnerUtils.h:70-74 (XsmmRunnerUtils.cpp:288-306) */
TPP_XSMM_EXPORT void xsmm_brgemm_invoke(int64_t dtype, int64_t addr,
                                        void *alignedPtrA, int64_t offsetA,
                                        void *alignedPtrB, int64_t offsetB,
                                        void *alignedPtrC, int64_t offsetC,
                                        int64_t numBatches);

/* replaces XsmmRunnerUtils.h:76-79 (XsmmRunnerUtils.cpp:363-383) */
TPP_XSMM_EXPORT void xsmm_fused_brgemm_invoke(
    int64_t dtype, int64_t addr, void *alignedPtrA, int64_t offsetA,
    void *alignedPtrB, int64_t offsetB, void *alignedPtrC, int64_t offsetC,
    void *alignedPtrD, int64_t offsetD, int64_t numBatches);

/* replaces XsmmRunnerUtils.h:81-83 (XsmmRunnerUtils.cpp:459-469): no-op. */
TPP_XSMM_EXPORT void xsmm_intel_amx_tile_config_invoke(int64_t dtype,
                                                       int64_t addr,
                                                       void *alignedPtrA,
                                                       int64_t offset);

/* ---- perf timers (runtime/PerfRunnerUtils.h:22-24, .cpp:23-35) -------------
 * Same wall-clock semantics; perf_stop_timer first drains every stream this
 * library launched on, because the timed region is a host loop of asynchronous
 * invokes (lib/TPP/Conversion/ConvertPerfToLoops/ConvertPerfToLoops.cpp:47-52). */
TPP_XSMM_EXPORT int64_t perf_start_timer(void);
TPP_XSMM_EXPORT double perf_stop_timer(int64_t startTimestamp);

/* The one libxsmm symbol the COMPILER links (VNNIUtils.cpp:36): VNNI blocking
 * factor for a datatype. Returns 2 for bf16 (the VNNI-2 layout is supported by
 * the BRGEMM kernels) unless TPP_XSMM_VNNI=0 is set in the environment, which
 * makes the compiler keep flat [K][N] weights (the TMA-native layout). */
TPP_XSMM_EXPORT int libxsmm_cpuid_dot_pack_factor(int datatype);

/* ---- CUDA-side extensions (new; the reference has no equivalent) ------------
 * They play the role of the reference GPU path's gpu.alloc/gpu.memcpy argument
 * offload (lib/TPP/Runner/MLIRBench.cpp:176-205): residency is explicit. */

/* Use this CUDA stream (a cudaStream_t / CUstream as void*) for all subsequent
 * invokes issued by the calling thread. NULL selects the per-thread default
 * stream of the library. */
TPP_XSMM_EXPORT void xsmm_cuda_set_stream(void *stream);
TPP_XSMM_EXPORT void *xsmm_cuda_get_stream(void);
/* Create / destroy a non-blocking CUDA stream without linking the CUDA runtime (JIT'd callers have only this
 * library). A thread that software-pipelines independent steps (upload of step i+1 under the kernels of step i under
 * the download of step i-1) creates one stream per in-flight step and switches with xsmm_cuda_set_stream; every
 * stream has its own kernel scratch, so work on different streams may overlap. destroy waits for the stream. */
TPP_XSMM_EXPORT void *xsmm_cuda_stream_create(void);
TPP_XSMM_EXPORT void xsmm_cuda_stream_destroy(void *stream);

/* Block until every invoke issued so far (all threads) has completed. */
TPP_XSMM_EXPORT void xsmm_cuda_sync(void);
/* Block until everything issued on the calling thread's stream has completed. */
TPP_XSMM_EXPORT void xsmm_cuda_stream_sync(void);

/* Register a host range: pins it and creates a device mirror of the same size.
 * Invokes whose operands lie inside a registered range run on the mirror with no
 * implicit copies. upload!=0 copies host->mirror now. Returns 0 on success. */
TPP_XSMM_EXPORT int64_t xsmm_cuda_register_host(void *host, int64_t bytes,
                                                int64_t upload);
TPP_XSMM_EXPORT int64_t xsmm_cuda_unregister_host(void *host);
/* Asynchronous (stream-ordered) host->mirror / mirror->host copies of a
 * sub-range of a registered range. */
TPP_XSMM_EXPORT int64_t xsmm_cuda_update_device(void *host, int64_t bytes);
TPP_XSMM_EXPORT int64_t xsmm_cuda_update_host(void *host, int64_t bytes);
/* Device address that mirrors a registered host address (NULL if none). */
/* Asynchronous forms for software-pipelined callers (independent steps, one buffer set per in-flight step).
 * upload_async copies host->mirror on a dedicated upload stream: it overlaps kernels already queued, and every invoke
 * (or graph launch) issued after the call waits for it. download_async copies mirror->host on a dedicated download
 * stream once the invokes issued before the call are done, and overlaps whatever is issued after it; wait_host blocks
 * until the latest download_async of that host address has landed (returns -1 if the calling thread has none on
 * record). The caller must not upload into a mirror that queued invokes still read, nor let invokes overwrite a
 * mirror whose download has not been waited for - i.e. consume step s before reusing its buffers for step s+depth.
 * Inside xsmm_cuda_graph_begin/end the copies become parallel branches of the captured graph (a graph that holds
 * several steps overlaps their copies with each other's kernels); such downloads are complete when the graph
 * launch is (xsmm_cuda_stream_sync). All return 0 on success, -1 if the range is not registered. */
TPP_XSMM_EXPORT int64_t xsmm_cuda_upload_async(void *host, int64_t bytes);
TPP_XSMM_EXPORT int64_t xsmm_cuda_download_async(void *host, int64_t bytes);
TPP_XSMM_EXPORT int64_t xsmm_cuda_wait_host(void *host);
TPP_XSMM_EXPORT void *xsmm_cuda_device_ptr(void *host);

/* CUDA-graph capture of a sequence of invokes (the body of a perf.bench loop):
 *   xsmm_cuda_graph_begin();  ...invokes on device / mirrored operands...;
 *   g = xsmm_cuda_graph_end();  then  xsmm_cuda_graph_launch(g) any number of times.
 * Replaying the graph re-runs exactly the captured kernels (same operands) with one
 * host call instead of one launch per invoke. Capture uses the calling thread's
 * stream (an internal one if the thread is on the legacy default stream). Plain host
 * operands cannot be captured (strict mode is synchronous): the invoke exit(-1)s.
 * graph_end returns 0 if the capture failed. */
TPP_XSMM_EXPORT int64_t xsmm_cuda_graph_begin(void);
TPP_XSMM_EXPORT int64_t xsmm_cuda_graph_end(void);
TPP_XSMM_EXPORT void xsmm_cuda_graph_launch(int64_t graph);
TPP_XSMM_EXPORT void xsmm_cuda_graph_destroy(int64_t graph);

/* Lazy mode (per thread; TPP_XSMM_LAZY=1 turns it on for every thread): outside a graph capture, BRGEMM and tile-move
 * invokes on device / registered operands are QUEUED instead of launched and go out - folded into layers, chained and
 * batched exactly as a captured sequence would be - at the next flush point: xsmm_cuda_sync, xsmm_cuda_stream_sync,
 * perf_start_timer / perf_stop_timer, the update / upload / download calls, xsmm_cuda_set_stream, xsmm_cuda_graph_*, an
 * invoke that cannot be queued (plain host operands, f32, ...), or 16384 queued invokes. An invoke loop without any
 * graph call (tools/tpp-run with device arguments, patches/0004) then runs on the fused kernels. Contract: drain through
 * one of those calls before operands are read, written or freed by any other means (cudaMemcpy, cudaFree, another
 * stream). Turning it off flushes. */
TPP_XSMM_EXPORT void xsmm_cuda_set_lazy(int64_t on);

/* Function-local temporaries. The reference's generated MLP keeps every layer's output except the last in a buffer that
 * lives and dies inside the kernel function (tools/mlir-gen/MLIRGen.cpp:255-261, 821-827: tensor.empty + fill inside
 * `entry`, memref.alloc / dealloc after bufferisation): nobody outside the function can observe it. A caller that marks such
 * a buffer (device pointer, or a host range registered with xsmm_cuda_register_host; the patched runner does it next to
 * the allocation) allows the runtime to treat its CONTENTS as dead once the last invoke that
 * reads them inside a fused launch has done so: the pair-per-chain kernel then drops those cache lines from L2
 * (discard.global.L2) instead of letting them be written back to HBM. Results of the invokes that consume the buffer are
 * unchanged; what a later read of the buffer itself returns is unspecified - so mark only buffers whose ONLY reader is the
 * invoke that directly follows their producer (the inter-layer activations of an MLP; not a buffer that also feeds a skip
 * connection or is read again by a later launch). Unmark before the memory is freed or reused as something observable.
 * Marks only take effect in launches captured / queued after the call, and a captured graph keeps the treatment it was
 * captured with: destroy graphs that were recorded while a buffer was marked before giving the buffer another role. */
TPP_XSMM_EXPORT void xsmm_cuda_mark_temporary(void *ptr, int64_t bytes);
TPP_XSMM_EXPORT void xsmm_cuda_unmark_temporary(void *ptr);

/* Introspection used by the tests and by bench.py's "gpu_launches". */
TPP_XSMM_EXPORT int64_t xsmm_cuda_launch_count(void);
/* Name of the kernel variant the last invoke on this thread launched
 * (e.g. "brgemm_tc_bf16_128x128x64"); points to static storage. */
TPP_XSMM_EXPORT const char *xsmm_cuda_last_kernel(void);
/* Name of the kernel variant a dispatch handle resolved to. */
TPP_XSMM_EXPORT const char *xsmm_cuda_handle_kernel(int64_t addr);
/* Debug: with TPP_XSMM_TC_TRACE=2 in the environment, print the wall-clock timeline of the
 * most recent BRGEMM launches (kernel overlap under PDL / graph replay) to stderr. */
TPP_XSMM_EXPORT void xsmm_cuda_debug_dump_trace(void);
/* Debug / test hook (no device needed): the hazard test the capture path uses before it adds a tile move to a
 * batched run - do two pitched rectangles of bytes (rows x width bytes, pitch ld bytes) share a byte? Exact for
 * equal pitches (tiles of one matrix interleave in address space without touching), conservative otherwise. */
TPP_XSMM_EXPORT int64_t xsmm_cuda_debug_rects_overlap(const void *a, int64_t a_rows, int64_t a_width, int64_t a_ld,
                                                      const void *b, int64_t b_rows, int64_t b_width, int64_t b_ld);
/* Debug / test hook (no device needed): the layer the capture path folds a run of `num` bf16 tile invokes into (one
 * descriptor, operand offsets in elements per invoke; d_off may be NULL). out[0..7] = grid_n, grid_k, a_step, b_step,
 * c_step_n, c_step_k, d_step, invokes folded into the first layer (1 = not a grid). */
TPP_XSMM_EXPORT int64_t xsmm_cuda_debug_fold_grid(int64_t m, int64_t n, int64_t k, int64_t lda, int64_t ldb, int64_t ldc,
                                                  int64_t stride_a, int64_t stride_b, int64_t flags, int64_t batch,
                                                  int64_t num, const int64_t *a_off, const int64_t *b_off,
                                                  const int64_t *c_off, const int64_t *d_off, int64_t *out);
/* Debug / test hook (no device needed): does a run of `num` tile moves (byte offsets of source / destination per move)
 * walk a regular grid, i.e. can it be one TMA-to-TMA copy (tile_grid.cu)? Returns 1 and out[0..5] = J (inner count), I,
 * in_inner, in_outer, out_inner, out_outer (byte steps), or 0. In xsmm_cuda_debug_fold_grid, bit 40 of `flags` makes the
 * invokes f32 (4-byte elements). */
TPP_XSMM_EXPORT int64_t xsmm_cuda_debug_tile_grid(int64_t num, const int64_t *in_off, const int64_t *out_off, int64_t *out);
/* ABI version of this header. */
TPP_XSMM_EXPORT int64_t xsmm_cuda_abi_version(void);

#ifdef __cplusplus
}
#endif

#endif /* TPP_XSMM_ABI_H */
