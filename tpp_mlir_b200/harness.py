"""Replay of the call sequences tpp-run's JIT-compiled ``main`` executes.

tpp-run / mlir-gen cannot be built here (no LLVM/MLIR), so this module is the
in-container caller of the C-ABI. It issues exactly the dispatch/invoke stream
the reference pipeline emits for its benchmark workloads:

* ``MlpReplay``: ``mlir-gen --kernel=const --bias --relu --float-type=bf16
  --batch=MB --layers=... --tiles=bn,bk,bc [--vnni=2]`` lowered through
  DefaultTppPasses - one ``xsmm_fused_brgemm_dispatch`` hoisted out of the loops
  and, per layer, one ``xsmm_fused_brgemm_invoke`` per (iN, iK) output block
  (tools/mlir-gen/MLIRGen.cpp:632-681, benchmarks/config/omp/mlir-bf16.json:37).
* ``bench_loop``: tpp-run's timing protocol (lib/TPP/Runner/TppRunnerWrapper.cpp
  :115-130, lib/TPP/Runner/MLIRBench.cpp:265-300): warm-up clamp(N/100,1,50)
  calls, then N timed calls between perf_start_timer / perf_stop_timer, mean
  seconds per call.

Layouts (block-packed, as mlir-gen emits them):
  input  [MB/bn][C/bc][bn][bc]     weight [K/bk][C/bc][bc][bk]
  bias   [K/bk][bk]                output [MB/bn][K/bk][bn][bk]
  VNNI weight: [K/bk][C/bc][bc/2][bk][2]
"""
from __future__ import annotations

from dataclasses import dataclass, field

from . import xsmm


@dataclass
class MlpConfig:
    batch: int = 256
    layers: tuple = (1024, 1024, 1024, 1024)  # sizes: input, hidden..., output
    tiles: tuple = (256, 1024, 1024)          # (bn, bk, bc); bk == bc for multi-layer nets
    dtype: int = xsmm.BF16
    vnni: bool = False
    bias: bool = True
    relu: bool = True

    def __post_init__(self):
        bn, bk, bc = self.tiles
        if self.batch % bn:
            raise ValueError("batch must be a multiple of the bn tile")
        for c, k in zip(self.layers[:-1], self.layers[1:]):
            if c % bc or k % bk:
                raise ValueError("layer sizes must be multiples of the bc / bk tiles")
        if len(self.layers) > 2 and bk != bc:
            raise ValueError("multi-layer mlir-gen nets need bk == bc (layer l+1 consumes layer l's blocks)")

    @property
    def num_layers(self) -> int:
        return len(self.layers) - 1

    def flops(self) -> int:
        """BENCH_TOTAL_FLOPS as mlir-gen counts it (MLIRGen.cpp:313-334): 2*M*N*K per
        matmul plus M*N for the bias add and M*N for the relu."""
        total = 0
        for c, k in zip(self.layers[:-1], self.layers[1:]):
            total += 2 * self.batch * c * k
            if self.bias:
                total += self.batch * k
            if self.relu:
                total += self.batch * k
        return total

    def matmul_flops(self) -> int:
        return sum(2 * self.batch * c * k for c, k in zip(self.layers[:-1], self.layers[1:]))


@dataclass
class MlpReplay:
    """Holds the dispatched handles and replays one forward pass per ``forward()``.

    ``weights[l]``, ``biases[l]`` and the activation buffers are tensor-likes in the
    block-packed layouts above (device tensors, or registered / plain host memory).
    """

    cfg: MlpConfig
    weights: list
    biases: list
    acts: list                       # len == num_layers + 1; acts[0] is the input
    handles: list = field(default_factory=list)
    tile_cfg: list = field(default_factory=list)

    def __post_init__(self):
        cfg = self.cfg
        bn, bk, bc = cfg.tiles
        gemm_flags = xsmm.GEMM_FLAG_BETA_0 | (xsmm.GEMM_FLAG_ROWMAJOR_B_VNNI if cfg.vnni else 0)
        if cfg.dtype == xsmm.BF16:
            # IntelAMXTileConfig insertion adds these to every bf16 brgemm
            # (lib/TPP/Transforms/IntelAMXTileConfig.cpp:32-139); ignored on the GPU.
            gemm_flags |= xsmm.GEMM_FLAG_NO_RESET_TILECONFIG | xsmm.GEMM_FLAG_NO_SETUP_TILECONFIG
        for _ in range(cfg.num_layers):
            # dispatches are hoisted out of the loops (LowLevelParallelization.cpp:55-63);
            # identical arguments return the identical handle
            if cfg.bias or cfg.relu:
                h = xsmm.fused_brgemm_dispatch(
                    cfg.dtype, bn, bk, bc, bc, bk, bk, bn * bc, bc * bk, gemm_flags,
                    xsmm.UNARY_FLAG_NONE, xsmm.UNARY_RELU if cfg.relu else xsmm.UNARY_NONE,
                    xsmm.BINARY_FLAG_BCAST_COL_IN_0 if cfg.bias else 0,
                    xsmm.BINARY_ADD if cfg.bias else xsmm.BINARY_NONE)
            else:
                h = xsmm.brgemm_dispatch(cfg.dtype, bn, bk, bc, bc, bk, bk, bn * bc, bc * bk, gemm_flags)
            self.handles.append(h)
            if cfg.dtype == xsmm.BF16:
                self.tile_cfg.append(xsmm.intel_amx_tile_config_dispatch(cfg.dtype, bn, bk, bc, bc, bk, bk, bn * bc,
                                                                         bc * bk, gemm_flags))

    def forward(self):
        cfg = self.cfg
        bn, bk, bc = cfg.tiles
        fused = cfg.bias or cfg.relu
        for l in range(cfg.num_layers):
            c, k = cfg.layers[l], cfg.layers[l + 1]
            src, dst, w, b, h = self.acts[l], self.acts[l + 1], self.weights[l], self.biases[l], self.handles[l]
            nb_c, nb_k = c // bc, k // bk
            for i_n in range(cfg.batch // bn):       # scf.parallel in the reference (OpenMP threads)
                for i_k in range(nb_k):
                    off_a = i_n * nb_c * bn * bc
                    off_b = i_k * nb_c * bc * bk
                    off_c = (i_n * nb_k + i_k) * bn * bk
                    if fused:
                        xsmm.fused_brgemm_invoke(cfg.dtype, h, src, off_a, w, off_b, dst, off_c,
                                                 b if cfg.bias else None, i_k * bk, nb_c)
                    else:
                        xsmm.brgemm_invoke(cfg.dtype, h, src, off_a, w, off_b, dst, off_c, nb_c)
        return self.acts[-1]

    @property
    def invokes_per_forward(self) -> int:
        cfg = self.cfg
        bn, bk, _ = cfg.tiles
        return sum((cfg.batch // bn) * (k // bk) for k in cfg.layers[1:])


def bench_loop(fn, n: int):
    """tpp-run's protocol: warm-up clamp(n/100,1,50) calls, n timed calls, mean seconds."""
    warm = min(max(n // 100, 1), 50)
    for _ in range(warm):
        fn()
    t0 = xsmm.perf_start_timer()
    for _ in range(n):
        fn()
    return xsmm.perf_stop_timer(t0) / n


# ---- tensor.pack / tensor.unpack as the reference lowers them: one unary TPP per tile -------------------
class PackReplay:
    """The invoke sequence of a tiled ``tensor.pack`` / ``tensor.unpack`` (inner_dims_pos = [0, 1], inner_tiles =
    [bm, bn], optional outer_dims_perm = [1, 0]): the reference tiles the op by one outer tile and turns every tile's
    copy into a ``xsmm.unary identity`` with ldi / ldo (lib/TPP/Transforms/LowerPacksAndUnpacks.cpp:143-250;
    benchmarks/mlir/fp32-pack-gemm-operand-a-512x1024.mlir). ``transpose_tiles`` replays the variant whose tiles are
    stored transposed ([..][bn][bm], ``xsmm.unary transpose`` per tile)."""

    def __init__(self, dtype, m, n, bm, bn, outer_perm=(0, 1), unpack=False, transpose_tiles=False):
        self.dtype, self.m, self.n, self.bm, self.bn = dtype, m, n, bm, bn
        self.outer_perm, self.unpack, self.transpose_tiles = tuple(outer_perm), unpack, transpose_tiles
        kind = xsmm.UNARY_TRANSPOSE if transpose_tiles else xsmm.UNARY_IDENTITY
        if not unpack:       # flat tile (pitch n) -> packed tile (contiguous)
            self.handle = xsmm.unary_dispatch(kind, dtype, bm, bn, n, bm if transpose_tiles else bn, 0)
        elif transpose_tiles:   # packed [bn][bm] tile -> flat [bm][bn] tile
            self.handle = xsmm.unary_dispatch(kind, dtype, bn, bm, bm, n, 0)
        else:
            self.handle = xsmm.unary_dispatch(kind, dtype, bm, bn, bn, n, 0)

    @property
    def num_tiles(self) -> int:
        return (self.m // self.bm) * (self.n // self.bn)

    def run(self, flat, packed) -> None:
        """pack: flat -> packed; unpack: packed -> flat. One xsmm_unary_invoke per tile."""
        mb, nb = self.m // self.bm, self.n // self.bn
        for i in range(mb):
            for j in range(nb):
                off_flat = i * self.bm * self.n + j * self.bn
                slot = j * mb + i if self.outer_perm == (1, 0) else i * nb + j
                off_packed = slot * self.bm * self.bn
                if self.unpack:
                    xsmm.unary_invoke(self.dtype, self.handle, packed, off_packed, flat, off_flat)
                else:
                    xsmm.unary_invoke(self.dtype, self.handle, flat, off_flat, packed, off_packed)


# ---- layout helpers (torch; used by tests / bench to build the packed operands) ---------
def pack_activation(x, bn, bc):
    """[MB][C] -> [MB/bn][C/bc][bn][bc]"""
    mb, c = x.shape
    return x.reshape(mb // bn, bn, c // bc, bc).permute(0, 2, 1, 3).contiguous()


def unpack_activation(xp):
    """[MB/bn][K/bk][bn][bk] -> [MB][K]"""
    nb, kb, bn, bk = xp.shape
    return xp.permute(0, 2, 1, 3).reshape(nb * bn, kb * bk).contiguous()


def pack_weight(w, bk, bc):
    """flat [C][K] (the matmul's B operand) -> [K/bk][C/bc][bc][bk]"""
    c, k = w.shape
    return w.reshape(c // bc, bc, k // bk, bk).permute(2, 0, 1, 3).contiguous()


def vnni_pack_weight(wp, factor=2):
    """[K/bk][C/bc][bc][bk] -> [K/bk][C/bc][bc/v][bk][v] (v = 2, or 4 for mlir-gen --vnni=4)"""
    kb, cb, bc, bk = wp.shape
    return wp.reshape(kb, cb, bc // factor, factor, bk).permute(0, 1, 2, 4, 3).contiguous()


# ---- native replay loop (csrc/harness/replay.cpp) ---------------------------------------------
import ctypes as _ct  # noqa: E402


class _TppMlpSet(_ct.Structure):
    _fields_ = [("acts", _ct.c_void_p * 9), ("weights", _ct.c_void_p * 8), ("biases", _ct.c_void_p * 8)]


_replay_lib = None


def replay_lib():
    """libtpp_replay.so: the hand-written equivalent of tpp-run's JIT-compiled hot loop. Pure host
    C++ that calls only the C-ABI; Python stays out of the timed loop."""
    global _replay_lib
    if _replay_lib is None:
        import os

        from . import _build

        xsmm.LIB  # the ABI library must be loaded (RTLD_GLOBAL) first
        path = _build.replay_path()
        if not os.path.exists(path):
            path = _build.build_replay()
        lib = _ct.CDLL(path)
        i64, p = _ct.c_int64, _ct.c_void_p
        lib.tpp_replay_mlp.argtypes = [i64, i64, p, p, i64, i64, i64, i64, p, i64, i64, i64, i64]
        lib.tpp_replay_mlp.restype = None
        lib.tpp_replay_mlp_e2e.argtypes = [i64, i64, p, p, i64, i64, i64, i64, p, i64, i64, i64, i64, p]
        lib.tpp_replay_mlp_e2e.restype = i64
        lib.tpp_replay_mlp_e2e_pipelined.argtypes = [i64, i64, p, p, i64, i64, i64, i64, p, i64, i64, i64, i64, i64, p, p]
        lib.tpp_replay_mlp_e2e_pipelined.restype = i64
        lib.tpp_replay_mlp_graph.argtypes = [i64, i64, p, p, i64, i64, i64, i64, p, i64, p, i64, i64, i64, i64]
        lib.tpp_replay_mlp_graph.restype = i64
        lib.tpp_replay_mlp_graph_unrolled.argtypes = [i64, i64, p, p, i64, i64, i64, i64, p, p, i64, i64, i64]
        lib.tpp_replay_mlp_graph_unrolled.restype = i64
        _replay_lib = lib
    return _replay_lib


class NativeMlpLoop:
    """K forward passes of an MlpReplay-style net issued from native code.

    ``sets`` is a list of (acts, weights, biases) tuples of tensor-likes; consecutive steps rotate
    through them (bench.py uses more bytes than the L2 holds)."""

    def __init__(self, cfg: MlpConfig, handles, sets):
        self.cfg = cfg
        n = cfg.num_layers
        assert n <= 8
        self._handles = (_ct.c_int64 * n)(*handles)
        self._sizes = (_ct.c_int64 * (n + 1))(*cfg.layers)
        self._sets = (_TppMlpSet * len(sets))()
        self._keep = sets
        for s, (acts, weights, biases) in zip(self._sets, sets):
            for i, a in enumerate(acts):
                s.acts[i] = xsmm._ptr(a)
            for i, w in enumerate(weights):
                s.weights[i] = xsmm._ptr(w)
            for i, b in enumerate(biases):
                s.biases[i] = xsmm._ptr(b)
        self.num_sets = len(sets)
        self._step = 0

    def reset(self) -> None:
        """Next run()/run_graph() starts at operand set 0 again (same graph decomposition for the same length)."""
        self._step = 0

    def run(self, steps: int) -> None:
        cfg = self.cfg
        bn, bk, bc = cfg.tiles
        replay_lib().tpp_replay_mlp(cfg.dtype, cfg.num_layers, self._handles, self._sizes, cfg.batch, bn, bk, bc,
                                    self._sets, self.num_sets, self._step, steps, 1 if cfg.bias else 0)
        self._step += steps

    def run_graph(self, steps: int, group: bool = True) -> None:
        """Like run(), but the invoke sequence is captured once into CUDA graphs (xsmm_cuda_graph_begin/end)
        and replayed: one graph per operand set, plus - with ``group`` - one graph holding a full rotation
        over all operand sets, so a graph launch covers ``num_sets`` forward passes."""
        cfg = self.cfg
        bn, bk, bc = cfg.tiles
        if not hasattr(self, "_graphs"):
            self._graphs = (_ct.c_int64 * (self.num_sets + 2))()   # per-set graphs, the rotation graph, the loop's id
        rc = replay_lib().tpp_replay_mlp_graph(cfg.dtype, cfg.num_layers, self._handles, self._sizes, cfg.batch, bn,
                                               bk, bc, self._sets, self.num_sets, self._graphs, self._step, steps,
                                               1 if cfg.bias else 0, 1 if group else 0)
        if rc != 0:
            raise RuntimeError("CUDA graph capture of the MLP forward failed")
        self._step += steps

    def run_graph_unrolled(self, steps: int, unroll: int = 16) -> None:
        """tpp-run's own benchmark loop: `steps` forward passes on operand set 0, back to back, replayed from a graph
        that holds `unroll` consecutive forward passes (the loop body unrolled before capture)."""
        cfg = self.cfg
        bn, bk, bc = cfg.tiles
        key = f"_unrolled_{unroll}"
        if not hasattr(self, key):
            setattr(self, key, (_ct.c_int64 * 2)())
        rc = replay_lib().tpp_replay_mlp_graph_unrolled(cfg.dtype, cfg.num_layers, self._handles, self._sizes, cfg.batch, bn,
                                                        bk, bc, self._sets, getattr(self, key), unroll, steps,
                                                        1 if cfg.bias else 0)
        if rc != 0:
            raise RuntimeError("CUDA graph capture of the unrolled MLP loop failed")

    def run_e2e(self, steps: int, elem_size: int = 2, graph: bool = True) -> None:
        """Host buffers registered with xsmm.register_host: per step H2D(input), layers, D2H(output), stream
        sync. graph=True replays the captured step (copies included) with one host call per step."""
        cfg = self.cfg
        bn, bk, bc = cfg.tiles
        if not hasattr(self, "_e2e_graph"):
            self._e2e_graph = _ct.c_int64(0)
        rc = replay_lib().tpp_replay_mlp_e2e(cfg.dtype, cfg.num_layers, self._handles, self._sizes, cfg.batch, bn, bk,
                                             bc, self._sets, steps, 1 if cfg.bias else 0, elem_size,
                                             1 if graph else 0, _ct.byref(self._e2e_graph))
        if rc != 0:
            raise RuntimeError("e2e replay failed (graph capture)")

    PIPE_MODES = {"async": 0, "grouped": 1, "streams": 2, "batch2": 4, "batch3": 5, "batch4": 6, "batch6": 8}

    @classmethod
    def pipe_mode(cls, mode: str) -> int:
        """"batchG" for any G >= 1: groups of G steps share one captured graph (mode code 2 + G)."""
        if mode in cls.PIPE_MODES:
            return cls.PIPE_MODES[mode]
        if mode.startswith("batch") and mode[5:].isdigit() and int(mode[5:]) >= 1:
            return 2 + int(mode[5:])
        raise KeyError(mode)

    def run_e2e_pipelined(self, steps: int, elem_size: int = 2, mode: str = "async") -> int:
        """Throughput form of run_e2e: every operand set is a pipeline slot, so uploads, kernels and downloads
        of neighbouring steps overlap; every step still moves its input and its output across PCIe.
        mode "async": xsmm_cuda_upload_async / graph replay / download_async per step, wait_host before a slot is
        reused; "grouped": one captured graph holds num_sets steps (copies are parallel branches);
        "streams": one stream + one captured step graph per slot; "batchG": groups of G steps share one captured graph
        of G independent layer chains (one interleaved chain-kernel launch), num_sets / G groups in flight.
        Returns the number of steps run."""
        cfg = self.cfg
        bn, bk, bc = cfg.tiles
        key = "_pipe_" + mode
        if not hasattr(self, key):
            setattr(self, key, ((_ct.c_int64 * (self.num_sets + 1))(), (_ct.c_void_p * self.num_sets)()))
        graphs, streams = getattr(self, key)
        rc = replay_lib().tpp_replay_mlp_e2e_pipelined(cfg.dtype, cfg.num_layers, self._handles, self._sizes, cfg.batch,
                                                       bn, bk, bc, self._sets, self.num_sets, steps,
                                                       1 if cfg.bias else 0, elem_size, self.pipe_mode(mode), graphs,
                                                       streams)
        if rc < 0:
            raise RuntimeError("pipelined e2e replay failed (graph capture)")
        return rc
