#!/usr/bin/env python
"""bench.py - fused-BRGEMM MLP (bf16, 3 x 1024^2, batch 256 per GPU) through the xsmm C-ABI.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward pass of the reference's benchmark workload
(`mlir-gen --kernel=const --bias --relu --float-type=bf16 --batch=256
--layers=1024,1024,1024,1024 --tiles=256,1024,1024`, benchmarks/config/omp/mlir-bf16.json:34-62
with the GPU tile setting of SURVEY.md Appendix B): 3 x xsmm_fused_brgemm_invoke
(m=256, n=1024, k=1024, bias add + ReLU fused), issued by the native replay loop
(tpp_mlir_b200/csrc/harness/replay.cpp) exactly as tpp-run's JIT-compiled loop would.

Metric = the reference's own: BENCH_TOTAL_FLOPS / mean seconds / 1e9 (benchmarks/harness/
controller.py:187-192), FLOPs counted as mlir-gen does (MLIRGen.cpp:313-334).

N > 1: the batch dimension is sharded (weak scaling: 256 rows per GPU, global batch 256*N,
BASELINE config 5 at N=8); weights/biases are broadcast once from rank 0 over NCCL; there is no
collective inside the timed loop. Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

LAYERS = (1024, 1024, 1024, 1024)
CHAIN_GROUP = 74        # forward passes per launch group: one per CTA pair of the pair-per-chain kernel (148 SMs / 2)
PIPE_GROUPS = int(os.environ.get("TPP_BENCH_PIPE_GROUPS", "3"))   # launch groups in flight in the end-to-end leg (upload | kernel | download)
BATCH_PER_GPU = 256
TILES = (256, 1024, 1024)
L2_BYTES = 126 * 1024 * 1024
METRIC = "fused_brgemm_mlp_bf16_3x1024_b256_gflops"
UNIT = "GFLOP/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained"), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0,
            "source": "fallback (B200_PROFILING.md)"}


# ---- clocks sampler (NVML; same counters nvidia-smi prints) -----------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.samples = []
        self._stop = threading.Event()
        self._thread = None
        self.max_mhz = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                self.samples.append((time.perf_counter(), mhz, reasons))
            except Exception:
                pass
            time.sleep(0.0005)

    def start(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=1.0)

    def summary(self, t0: float, t1: float):
        if self._nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        inside = [s for s in self.samples if t0 <= s[0] <= t1] or self.samples[-3:]
        mhz = sorted(s[1] for s in inside)
        bits = 0
        for s in inside:
            bits |= s[2]
        reasons = [name for bit, name in self.REASONS.items() if bits & bit]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                "samples": len(inside)}


# ---- workload ------------------------------------------------------------------------------------
def make_host_data(seed=123):
    """TensorInit 'normal' seed 123 (tpp-run --seed 123 --splat-to-random --init-type normal): splat
    constants first in op order (W1,b1,W2,b2,W3,b3), then the kernel argument (the input)."""
    import oracle

    gen = oracle.TensorInit("normal", oracle.BF16, seed)
    Ws, bs = [], []
    for c, k in zip(LAYERS[:-1], LAYERS[1:]):
        Ws.append(gen.fill(c, k))
        bs.append(gen.fill(k))
    return gen, Ws, bs


def capture_stdout():
    """Send everything written to fd 1 from here on (NCCL's version banner, library chatter) to stderr; returns the
    saved descriptor for restore_stdout(). Rank 0's JSON line must be the ONLY line on stdout."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def restore_stdout(saved):
    sys.stdout.flush()
    os.dup2(saved, 1)
    os.close(saved)


def h_out_bits(h_acts, harness, bn, bk):
    """first 8 rows of a host-side output buffer (block-packed) as int32 bf16 bit patterns"""
    import numpy as np

    o = harness.unpack_activation(h_acts[-1].reshape(BATCH_PER_GPU // bn, LAYERS[-1] // bk, bn, bk))[:8].numpy()
    return o.view(np.uint16).astype(np.int32)


def cpu_arm(steps, warmup, total_budget_s, verbose=False):
    """The reference's CPU path for this workload, timed on this box's host cores.
    kind = "port": oracle/ (libxsmm is not buildable offline). Each step is a
    bounded sample: `rows` of the 256 batch rows through all three layers."""
    import numpy as np

    import oracle

    oracle.use_native(True)  # -march=native build for the box it is timed on
    oracle.lib()
    gen, Ws, bs = make_host_data()
    x = gen.fill(BATCH_PER_GPU, LAYERS[0])

    def forward(rows):
        a = x[:rows]
        for W, b in zip(Ws, bs):
            y = np.empty((rows, W.shape[1]), np.uint16)
            if not oracle.fused_brgemm_fast(2, rows, W.shape[1], W.shape[0], W.shape[0], W.shape[1], W.shape[1], 0, 0, 4,
                                            5, 4, 1, a, W, y, b, 1):
                oracle.fused_brgemm(2, rows, W.shape[1], W.shape[0], W.shape[0], W.shape[1], W.shape[1], 0, 0, 4, 0, 5,
                                    4, 1, a, W, y, b, 1)
            a = y
        return a

    # "all the host threads it can use": the usable count is not os.cpu_count() inside a container with a CPU
    # quota - try a ladder of thread counts on one forward pass each and keep the fastest
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ladder = sorted({t for t in (4, 8, 16, 32, 64, 96, 128, avail) if t <= avail} | {min(avail, 8)})
    best_t, t1 = 1, float("inf")
    for nt in ladder:
        oracle.set_num_threads(nt)
        forward(BATCH_PER_GPU)  # first touch / thread pool spin-up
        t = time.perf_counter()
        forward(BATCH_PER_GPU)
        dt1 = time.perf_counter() - t
        if dt1 < t1:
            best_t, t1 = nt, dt1
    oracle.set_num_threads(best_t)
    rows = BATCH_PER_GPU
    n_calls = steps + warmup
    if t1 * n_calls > total_budget_s:
        rows = int(BATCH_PER_GPU * total_budget_s / (t1 * n_calls)) // 8 * 8
        rows = max(8, min(BATCH_PER_GPU, rows))
    for _ in range(warmup):
        forward(rows)
    t = time.perf_counter()
    for _ in range(steps):
        forward(rows)
    dt = (time.perf_counter() - t) / steps
    flops = sum(2 * rows * c * k + 2 * rows * k for c, k in zip(LAYERS[:-1], LAYERS[1:]))
    return {"value": flops / dt / 1e9, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
            "sample": f"{rows} of {BATCH_PER_GPU} batch rows x 3 layers per step, {steps} steps "
                      f"(oracle/xsmm_oracle_fast.c: {oracle.fast_isa()}, OpenMP {oracle.num_threads()} threads, "
                      f"gcc -O3 -march=native; libxsmm itself is not buildable offline)",
            "ms_per_step": dt * 1e3, "rows": rows}


def reference_main(args, rank):
    if rank != 0:
        return 0
    res = cpu_arm(args.steps, args.warmup, total_budget_s=150.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic (TensorInit normal, seed 123)",
        "config": {"workload": "fused_brgemm MLP 3x(256x1024x1024)+bias+relu bf16, batch 256", "layers": list(LAYERS),
                   "global_batch": BATCH_PER_GPU, "parallelism": "host cores (OpenMP)"},
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=14800)   # 100 rotations of 148 operand sets
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="graph", choices=["graph", "direct"],
                    help="graph: replay the captured invoke sequence (one host call per step); "
                         "direct: one C-ABI invoke per layer per step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return reference_main(args, rank)

    import numpy as np
    import torch
    import torch.distributed as dist

    from tpp_mlir_b200 import harness, shard, xsmm

    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: this backend has no CPU path"}))
        return 1
    saved_stdout = capture_stdout()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    n_gpus = world

    bn, bk, bc = TILES
    cfg = harness.MlpConfig(batch=BATCH_PER_GPU, layers=LAYERS, tiles=TILES)

    # ---- data: rank 0 generates, NCCL broadcasts weights/biases; each rank owns its batch shard ----
    def to_dev(a):
        return torch.from_numpy(a.view(np.int16)).to(dev)

    if rank == 0:
        gen, Ws, bs = make_host_data()
        x_all = gen.fill(BATCH_PER_GPU * n_gpus, LAYERS[0])
        w_dev = [harness.pack_weight(to_dev(W), bk, bc) for W in Ws]
        b_dev = [to_dev(b) for b in bs]
        x_dev_all = to_dev(x_all)
    else:
        w_dev = [torch.empty(k // bk, c // bc, bc, bk, dtype=torch.int16, device=dev)
                 for c, k in zip(LAYERS[:-1], LAYERS[1:])]
        b_dev = [torch.empty(k, dtype=torch.int16, device=dev) for k in LAYERS[1:]]
        x_dev_all = torch.empty(BATCH_PER_GPU * n_gpus, LAYERS[0], dtype=torch.int16, device=dev)
    # the one collective of this path: parameters (and the synthetic input), once, outside the timed loop
    shard.broadcast_parameters(w_dev + b_dev + [x_dev_all], src=0)
    lo, hi = shard.shard_bounds(BATCH_PER_GPU * n_gpus, rank, n_gpus, tile_m=bn)
    x_shard = x_dev_all[lo:hi].contiguous()
    x_packed = harness.pack_activation(x_shard, bn, bc)

    # ---- rotate more bytes than the L2 holds so every step streams its operands from HBM ----
    set_bytes = sum(w.numel() * 2 for w in w_dev) + sum(b.numel() * 2 for b in b_dev) + 4 * BATCH_PER_GPU * 1024 * 2
    # ... and at least two forward passes per CTA pair in one rotation graph, so that the pair-per-chain kernel
    # (one pair of SMs per forward pass) has two rounds of work per launch
    num_sets = max(L2_BYTES // set_bytes + 2, 2 * CHAIN_GROUP)
    sets = []
    for s in range(num_sets):
        acts = [x_packed.clone()] + [torch.zeros(BATCH_PER_GPU * k, dtype=torch.int16, device=dev) for k in LAYERS[1:]]
        sets.append((acts, [w.clone() for w in w_dev], [b.clone() for b in b_dev]))
    replay = harness.MlpReplay(cfg, sets[0][1], sets[0][2], sets[0][0])  # dispatches (hoisted, once)
    loop = harness.NativeMlpLoop(cfg, replay.handles, sets)
    stream = torch.cuda.current_stream(dev)
    xsmm.set_stream(stream.cuda_stream)

    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident number ("value") --------------------------------------------------------
    # mode "graph": every operand set's forward pass (3 invokes) is captured once through
    # xsmm_cuda_graph_begin/end and replayed - one host call per step; mode "direct": one
    # xsmm_fused_brgemm_invoke (one cudaLaunchKernelEx) per layer per step.
    run = loop.run_graph if args.mode == "graph" else loop.run
    run(max(args.warmup, num_sets))   # warm-up also captures the rotation graph
    loop.reset()
    run(args.steps)                   # rehearsal of the timed call: captures the graph of its partial last rotation
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    loop.reset()
    run(2 * num_sets)  # keep the GPU busy while the sampler gets going
    barrier()
    loop.reset()
    launches0 = xsmm.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    run(args.steps)
    t_issue = time.perf_counter() - t_wall0   # host time to issue all launches (no sync yet)
    ev1.record(stream)
    barrier()
    t_wall1 = time.perf_counter()
    launches = xsmm.launch_count() - launches0
    loop.reset()
    run(num_sets)                       # exactly one rotation: the name of the kernel the timed loop is made of
    timed_kernel = xsmm.last_kernel()   # (graph mode: the fused multi-chain kernel; leftover steps run single chains)
    sampler.stop()
    ms = ev0.elapsed_time(ev1)
    ms_max = shard.max_over_ranks(ms, device=dev)
    ms_per_step = ms_max / args.steps
    flops_step_rank = cfg.flops()
    value = flops_step_rank * n_gpus / (ms_per_step * 1e-3) / 1e9

    if os.environ.get("TPP_XSMM_TC_TRACE") in ("2", "3"):
        xsmm.LIB.xsmm_cuda_debug_dump_trace()

    # hot-L2 variant (what tpp-run measures: the same buffers every iteration), for information
    hot = harness.NativeMlpLoop(cfg, replay.handles, sets[:1])
    run_hot = hot.run_graph if args.mode == "graph" else hot.run
    run_hot(args.warmup)
    barrier()
    ev0.record(stream)
    run_hot(args.steps)
    ev1.record(stream)
    barrier()
    ms_hot = ev0.elapsed_time(ev1) / args.steps

    # the other issue mode, for information (same kernels, different host path)
    other = loop.run if args.mode == "graph" else loop.run_graph
    other_steps = min(args.steps, 500)
    other(max(args.warmup, num_sets))
    barrier()
    ev0.record(stream)
    other(other_steps)
    ev1.record(stream)
    barrier()
    ms_other = ev0.elapsed_time(ev1) / other_steps

    # ---- parity of what was just timed (rank-local, against the oracle on a row sample) ----------
    import oracle

    out = harness.unpack_activation(sets[0][0][-1].reshape(BATCH_PER_GPU // bn, LAYERS[-1] // bk, bn, bk))
    got = oracle.bf16_to_f32(out[:8].cpu().numpy().view(np.uint16))
    if rank == 0:
        a = x_all[:8]
        for W, b in zip(Ws, bs):
            y = np.empty((8, W.shape[1]), np.uint16)
            oracle.fused_brgemm(2, 8, W.shape[1], W.shape[0], W.shape[0], W.shape[1], W.shape[1], 0, 0, 4, 0, 5, 4, 1,
                                a, W, y, b, 1)
            a = y
        want = oracle.bf16_to_f32(a)
        rel = float(np.abs(got - want).max() / max(np.abs(want).max(), 1e-30))
    else:
        rel = None

    # ---- end-to-end through the C-ABI with HOST buffers (rank-local), H2D + D2H inside the timing ----
    host_sets = None
    e2e_steps = min(args.steps, 500)
    h_w = [w.cpu().contiguous().pin_memory() for w in w_dev]
    h_b = [b.cpu().contiguous().pin_memory() for b in b_dev]
    h_acts = [x_packed.cpu().contiguous().pin_memory()] + [torch.zeros(BATCH_PER_GPU * k, dtype=torch.int16).pin_memory()
                                                           for k in LAYERS[1:]]
    for tns in h_w + h_b + h_acts:
        xsmm.register_host(tns, upload=True)  # parameters + activation buffers mirrored once (like gpu.alloc)
    host_sets = [(h_acts, h_w, h_b)]
    e2e_loop = harness.NativeMlpLoop(cfg, replay.handles, host_sets)
    e2e_loop.run_e2e(max(3, args.warmup // 5))
    barrier()
    t0 = time.perf_counter()
    e2e_loop.run_e2e(e2e_steps)
    torch.cuda.synchronize(dev)
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    e2e_s = shard.max_over_ranks(e2e_s, device=dev)
    e2e_value = flops_step_rank * n_gpus / e2e_s / 1e9
    e2e_out = oracle.bf16_to_f32(harness.unpack_activation(
        h_acts[-1].reshape(BATCH_PER_GPU // bn, LAYERS[-1] // bk, bn, bk))[:8].numpy().view(np.uint16))
    # the single-step graph runs the full-K pass kernel, the device-timed loop the pair kernel: same math, the f32
    # summation order inside the tensor core may differ -> compare in bf16 ulps, and against the oracle below
    sync_ulp = int(np.abs(h_out_bits(h_acts, harness, bn, bk) - out[:8].cpu().numpy().view(np.uint16).astype(np.int32)).max())
    e2e_rel = float(np.abs(e2e_out - want).max() / max(np.abs(want).max(), 1e-30)) if rank == 0 else None
    # throughput form: PIPE_GROUPS groups of CHAIN_GROUP independent steps in flight (one buffer set per step); every
    # step still uploads its 512 KiB input and downloads its 512 KiB output
    depth = PIPE_GROUPS * CHAIN_GROUP
    # per group ONE pinned, registered host block per activation level; a step's buffers are slices of it, so that a
    # group's inputs (outputs) cross PCIe as one 37 MiB copy
    blocks, slot_acts = [], []
    xp_host = x_packed.cpu().contiguous().reshape(-1)
    for _ in range(PIPE_GROUPS):
        lvl = [torch.zeros(CHAIN_GROUP, BATCH_PER_GPU * k, dtype=torch.int16).pin_memory() for k in LAYERS]
        lvl[0][:] = xp_host
        for tns in lvl:
            xsmm.register_host(tns, upload=True)
        blocks += lvl
        for j in range(CHAIN_GROUP):
            slot_acts.append([b[j] for b in lvl])
    pipe_loop = harness.NativeMlpLoop(cfg, replay.handles, [(a, h_w, h_b) for a in slot_acts])
    pipe_mode = f"batch{CHAIN_GROUP}"
    # enough group iterations that filling and draining the 3-deep pipeline (one upload + kernel + download = ~1.5 ms,
    # inside the timing) does not dominate: 8 ... 24 rounds over the groups
    pipe_steps = min(max(args.steps // depth, 8), 24) * depth
    pipe_loop.run_e2e_pipelined(depth, mode=pipe_mode)
    torch.cuda.synchronize(dev)
    for a in slot_acts:
        a[-1].zero_()
    pipe_runs = []
    for _ in range(3):   # three timed repetitions; the median is reported
        barrier()
        t0 = time.perf_counter()
        ran = pipe_loop.run_e2e_pipelined(pipe_steps, mode=pipe_mode)   # returns after the last output reached the host
        pipe_runs.append(shard.max_over_ranks((time.perf_counter() - t0) / ran, device=dev))
    pipe_s = sorted(pipe_runs)[1]
    pipe_value = flops_step_rank * n_gpus / pipe_s / 1e9
    pipe_kernel = xsmm.last_kernel()
    e2e_ulp = 0
    got_bits = out[:8].cpu().numpy().view(np.uint16).astype(np.int32)
    for a in slot_acts:
        o = harness.unpack_activation(a[-1].reshape(BATCH_PER_GPU // bn, LAYERS[-1] // bk, bn, bk))[:8].numpy()
        e2e_ulp = max(e2e_ulp, int(np.abs(o.view(np.uint16).astype(np.int32) - got_bits).max()))
    for tns in h_w + h_b + h_acts + blocks:
        xsmm.unregister_host(tns)

    if rank != 0:
        if n_gpus > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0

    pk = peaks()
    # algorithmic HBM bytes of one forward pass: 3 weight matrices + 3 biases + input + output (intermediates stay in L2)
    set_bytes_algo = sum(c * k * 2 + k * 2 for c, k in zip(LAYERS[:-1], LAYERS[1:])) + 2 * BATCH_PER_GPU * 1024 * 2
    launches_per_step = launches / args.steps
    flops_per_launch = flops_step_rank / launches_per_step
    avg_launch_s = (ms / args.steps) * 1e-3 / launches_per_step
    achieved_tflops = flops_per_launch / avg_launch_s / 1e12
    bytes_per_launch = set_bytes_algo / launches_per_step
    traffic = None
    prof = os.path.join(ROOT, "profiles", "dominant_kernel.json")
    if os.path.exists(prof):
        with open(prof) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    cpu = None
    if n_gpus == 1 and not args.no_cpu_baseline:
        c = cpu_arm(steps=10, warmup=2, total_budget_s=20.0)
        cpu = {k: c[k] for k in ("value", "unit", "cores", "kind", "sample")}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic (TensorInit normal, seed 123; random-init weights)",
        "config": {"workload": "fused_brgemm MLP 3x(256x1024x1024)+bias+relu bf16, batch 256 per GPU "
                               "(BASELINE configs[2]; configs[4] at 8 GPUs)",
                   "layers": list(LAYERS), "tiles": list(TILES), "global_batch": BATCH_PER_GPU * n_gpus,
                   "parallelism": f"batch-sharded x{n_gpus}, weights broadcast once over NCCL",
                   "l2": f"rotating {num_sets} operand sets ({num_sets * set_bytes >> 20} MiB > 126 MiB L2), "
                         "inputs larger than L2",
                   "timing": "CUDA events on the launch stream, max over ranks",
                   "issue_mode": ("CUDA graph replay of the captured xsmm invoke sequence (xsmm_cuda_graph_*): one graph = "
                                  f"one rotation of {num_sets} forward passes on {num_sets} operand sets ({3 * num_sets} "
                                  "xsmm_fused_brgemm_invoke calls), which the runtime turns into ONE launch of the "
                                  "pair-per-chain kernel: each forward pass (a chain of 3 dependent layers) runs on one "
                                  "pair of SMs, 74 forward passes side by side; latency of a lone forward pass: "
                                  "extra.ms_per_step_single_forward"
                                  if args.mode == "graph" else "one xsmm_fused_brgemm_invoke per layer"),
                   "flops_per_step": flops_step_rank * n_gpus, "matmul_flops_per_step": cfg.matmul_flops() * n_gpus},
        "clocks": sampler.summary(t_wall0, t_wall1),
        "e2e": {"value": pipe_value, "unit": UNIT, "h2d_bytes_per_step": BATCH_PER_GPU * LAYERS[0] * 2,
                "d2h_bytes_per_step": BATCH_PER_GPU * LAYERS[-1] * 2, "ms_per_step": pipe_s * 1e3, "steps": pipe_steps,
                "path": "xsmm C-ABI on registered pinned host buffers, every step: xsmm_cuda_upload_async(input 512 KiB) "
                        "-> 3 xsmm_fused_brgemm_invoke (replayed from the captured sequence) -> xsmm_cuda_download_async("
                        f"output 512 KiB); steps are issued in groups of {CHAIN_GROUP} (one captured graph = one launch of the "
                        f"pair-per-chain kernel per group; the group's {CHAIN_GROUP} inputs / outputs are slices of one registered "
                        f"host block and move as one copy), {PIPE_GROUPS} groups in flight, xsmm_cuda_wait_host(output) before "
                        "a group's buffers are reused, so uploads, kernels and downloads of neighbouring groups overlap; wall "
                        "clock incl. the final drain; median of 3 repetitions; bound by PCIe (1 MiB per step; "
                        "scripts/pcie_probe.py: ~50 GB/s per direction with both directions busy = 10.5 us per step)",
                "pipeline_depth": depth, "kernel": pipe_kernel,
                "ms_per_step_repetitions": [t * 1e3 for t in pipe_runs],
                "max_ulp_diff_vs_device_run": e2e_ulp,
                "synchronous": {"value": e2e_value, "ms_per_step": e2e_s * 1e3,
                                "path": "same step, one in flight: graph launch -> stream sync, every step",
                                "rel_err_vs_oracle": e2e_rel, "max_ulp_diff_vs_device_run": sync_ulp}},
        "gpu_launches": launches,
        # the dominant kernel is HBM-bound (ncu: DRAM 5.3 TB/s busy, tensor pipe < 50 % of cycles): the roofline is the
        # measured copy bandwidth; the tensor-core view of the same launch is kept beside it
        "roofline": {"bound": "hbm", "achieved": bytes_per_launch / avg_launch_s / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                     "frac": bytes_per_launch / avg_launch_s / 1e9 / pk["hbm_gbs"], "traffic": traffic,
                     "kernel": timed_kernel, "peak_source": pk["source"],
                     "algorithmic_bytes_per_launch": bytes_per_launch, "avg_launch_us": avg_launch_s * 1e6,
                     "forward_passes_per_launch": args.steps / max(launches, 1),
                     "bytes_per_forward_pass": set_bytes_algo,
                     "tensor": {"achieved": achieved_tflops, "unit": "TFLOP/s", "peak_burst": pk["bf16_tflops"],
                                "frac_of_burst_peak": achieved_tflops / pk["bf16_tflops"],
                                "peak_sustained": pk["bf16_tflops_sustained"],
                                "frac_of_sustained_peak": (achieved_tflops / pk["bf16_tflops_sustained"]
                                                           if pk["bf16_tflops_sustained"] else None),
                                "flops_per_launch": flops_per_launch},
                     "note": "arithmetic intensity 219 FLOP/B (1.61 GFLOP over 7.35 MB: 3 weight matrices, biases, input, "
                             "output; intermediates stay in L2) is below the measured machine balance (1641 TF/s / 6.55 TB/s "
                             "= 251): with operand sets rotating through > L2 every weight byte comes from HBM and HBM "
                             "bounds the step (DESIGN.md 4.1d). traffic = DRAM bytes of one launch under ncu "
                             "(profiles/ncu_mlp_chain_pair_r1.json)"},
        "cpu_baseline": cpu,
        "extra": {"ms_per_step_single_forward": ms_hot,
                  "single_forward_note": "one operand set replayed back to back (L2-hot, what tpp-run itself measures): "
                                         "one forward pass per launch, the full-K pass kernel on 128 SMs",
                  "host_issue_us_per_launch": t_issue / max(launches, 1) * 1e6,
                  ("ms_per_step_direct_invokes" if args.mode == "graph" else "ms_per_step_graph_replay"): ms_other,
                  "parity_rel_err_vs_oracle": rel, "kernel": timed_kernel,
                  "per_layer_kernel": xsmm.handle_kernel(replay.handles[0])},
    }
    restore_stdout(saved_stdout)
    print(json.dumps(line))
    sys.stdout.flush()
    if n_gpus > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
