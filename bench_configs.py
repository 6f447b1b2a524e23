#!/usr/bin/env python
"""bench_configs.py - the BASELINE.json configurations that are NOT the headline line of bench.py,
measured through the same C-ABI on one B200, each against the roofline that bounds it:

  cfg2   single BRGEMM bf16 M=N=K=1024, batch 16 (34.36 GFLOP)            -> tensor peak
  cfg3b  MLP layer as the strided view k=64 x batch 16 (same math as cfg3)  -> tensor peak
  cfg4   unary vnni_2 4096x4096 bf16 (+ inverse, transpose, relu, bias add) -> HBM bandwidth
  cfg5   MLP 3x1024^2 at batch 2048 on one GPU (what 8 GPUs shard)          -> tensor peak
  pack   tensor.pack / unpack as per-tile unary TPPs, batched under capture   -> HBM bandwidth

Writes one JSON line per config to stdout (and to --out). Operands rotate over more bytes than the
126 MiB L2; time = CUDA events on the launch stream after warm-up.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from bench import peaks  # noqa: E402
from tpp_mlir_b200 import harness, xsmm  # noqa: E402

L2 = 126 << 20
BF16 = xsmm.BF16


def timed(launchers, iters, warmup=5):
    """launchers: list of zero-arg callables (one per rotating operand set)."""
    stream = torch.cuda.current_stream()
    xsmm.set_stream(stream.cuda_stream)
    for i in range(warmup):
        launchers[i % len(launchers)]()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(iters):
        launchers[i % len(launchers)]()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / iters


def rnd(*shape):
    return (torch.rand(*shape, device="cuda") * 0.5).to(torch.bfloat16)


def sets_needed(bytes_per_set):
    return L2 // bytes_per_set + 2


def cfg2(pk):
    m = n = k = 1024
    batch = 16
    h = xsmm.brgemm_dispatch(BF16, m, n, k, k, n, n, m * k, k * n, 4 | 64 | 128)
    ns = sets_needed(2 * batch * m * k * 2 + m * n * 2)
    S = [(rnd(batch, m, k), rnd(batch, k, n), torch.empty(m, n, dtype=torch.bfloat16, device="cuda")) for _ in range(ns)]
    fns = [lambda A=A, B=B, C=C: xsmm.LIB.xsmm_brgemm_invoke(BF16, h, A.data_ptr(), 0, B.data_ptr(), 0, C.data_ptr(), 0, batch)
           for A, B, C in S]
    t = timed(fns, 40)
    # parity on a slab: rows 0..7 against a float64 reduction of the same bf16 data
    A, B, C = S[0]
    fns[0]()
    torch.cuda.synchronize()
    want = torch.einsum("bik,bkj->ij", A[:, :8].double(), B.double())
    rel = float(((C[:8].double() - want).abs().max() / want.abs().max()).item())
    flops = 2.0 * m * n * k * batch
    kernel = xsmm.last_kernel()
    # the vendor library on the same arithmetic, for orientation only (not on any path of this repo): the batch reduction
    # folded into K, A' = [m][batch * k], B' = [batch * k][n], same rotating operand sets
    lib = None
    try:
        A2 = [A.permute(1, 0, 2).reshape(m, batch * k).contiguous() for A, _, _ in S]
        B2 = [B.reshape(batch * k, n) for _, B, _ in S]
        t_lib = timed([lambda a=a, b=b: torch.matmul(a, b) for a, b in zip(A2, B2)], 40)
        lib = {"what": "torch.matmul (cuBLAS) bf16 [1024 x 16384] x [16384 x 1024], same data", "seconds": t_lib,
               "gflops": flops / t_lib / 1e9}
        del A2, B2
    except Exception as e:
        lib = {"error": repr(e)}
    return {"config": "cfg2 brgemm bf16 1024x1024x1024 batch 16", "kernel": kernel, "seconds": t,
            "gflops": flops / t / 1e9, "vendor_library_same_shape": lib,
            "roofline": {"bound": "tensor", "achieved": flops / t / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": flops / t / 1e12 / pk["bf16_tflops"], "flops_per_launch": flops,
                         "algorithmic_bytes": 2 * batch * m * k * 2 + m * n * 2},
            "rotating_sets": ns, "rel_err_vs_f64": rel}


def mlp(pk, batch, name, tiles):
    layers = (1024,) * 4
    cfg = harness.MlpConfig(batch=batch, layers=layers, tiles=tiles)
    bn, bk, bc = tiles
    set_bytes = 3 * (1024 * 1024 * 2) + 4 * batch * 1024 * 2
    ns = sets_needed(set_bytes)
    sets = []
    for _ in range(ns):
        acts = [rnd(batch, 1024)] + [torch.empty(batch, 1024, dtype=torch.bfloat16, device="cuda") for _ in range(3)]
        sets.append((acts, [rnd(1024, 1024) * 0.1 for _ in range(3)], [rnd(1024) for _ in range(3)]))
    rp = harness.MlpReplay(cfg, sets[0][1], sets[0][2], sets[0][0])
    loop = harness.NativeMlpLoop(cfg, rp.handles, sets)
    stream = torch.cuda.current_stream()
    xsmm.set_stream(stream.cuda_stream)
    loop.run_graph(2 * ns)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    steps = 20 * ns
    e0.record(stream)
    loop.run_graph(steps)
    e1.record(stream)
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / steps
    return {"config": name, "kernel": xsmm.last_kernel(), "seconds_per_forward": t,
            "gflops": cfg.flops() / t / 1e9,
            "roofline": {"bound": "tensor", "achieved": cfg.flops() / t / 1e12, "peak": pk["bf16_tflops"],
                         "unit": "TFLOP/s", "frac": cfg.flops() / t / 1e12 / pk["bf16_tflops"]},
            "rotating_sets": ns, "launches_per_forward": rp.invokes_per_forward}


def cfg3b(pk):
    """One MLP layer issued as the strided BRGEMM view k=64 x batch 16 (lda=1024, stride_a=64, stride_b=65536)."""
    m, n = 256, 1024
    h = xsmm.fused_brgemm_dispatch(BF16, m, n, 64, 1024, 1024, 1024, 64, 64 * 1024, 4 | 64 | 128, 0, 5, 4, 1)
    ns = sets_needed(2 * 1024 * 1024 + 2 * m * 1024 * 2)
    S = [(rnd(m, 1024), rnd(1024, 1024), torch.empty(m, n, dtype=torch.bfloat16, device="cuda"), rnd(1024)) for _ in range(ns)]
    fns = [lambda A=A, B=B, C=C, D=D: xsmm.LIB.xsmm_fused_brgemm_invoke(BF16, h, A.data_ptr(), 0, B.data_ptr(), 0,
                                                                         C.data_ptr(), 0, D.data_ptr(), 0, 16)
           for A, B, C, D in S]
    t = timed(fns, 400, warmup=ns)
    flops = 2.0 * m * n * 1024 + 2 * m * n
    return {"config": "cfg3b fused_brgemm 256x1024 k=64 x batch 16 (strided view)", "kernel": xsmm.last_kernel(),
            "seconds": t, "gflops": flops / t / 1e9,
            "roofline": {"bound": "tensor", "achieved": flops / t / 1e12, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                         "frac": flops / t / 1e12 / pk["bf16_tflops"]}}


def cfg4(pk):
    """BASELINE configs[3]: unary vnni_2 4096x4096 bf16 and its inverse, rotating operands > L2."""
    rows = eltwise(pk, only=("cfg4 unary vnni_2 pack", "cfg4 inverse"))
    head = dict(rows[0])
    head["inverse"] = rows[1]
    return head


def reference_stream(pk, tiles=(32, 32, 32), vnni=True):
    """The reference's DEFAULT benchmark call stream (benchmarks/config/omp/mlir-bf16.json:37: --tiles=32,32,32 --vnni=2):
    768 xsmm_fused_brgemm_invoke calls of a 32x32x32 x batch-32 VNNI-2 BRGEMM per forward pass on block-packed operands,
    captured and replayed like bench.py's headline. Reports the throughput of that stream and which kernel ran it."""
    import numpy as np

    import bench

    gen, Ws, bs = bench.make_host_data()
    x = gen.fill(256, 1024)
    dev = torch.device("cuda", torch.cuda.current_device())

    def to_dev(a):
        return torch.from_numpy(a.view(np.int16)).to(dev)

    stream = torch.cuda.current_stream()
    xsmm.set_stream(stream.cuda_stream)
    probe = bench.MlpWorkload(256, to_dev(x), [to_dev(W) for W in Ws], [to_dev(b) for b in bs], tiles=tiles,
                              vnni=vnni, max_sets=1)
    probe.rotations(1)
    torch.cuda.synchronize()
    fused = "pair" in xsmm.last_kernel() or "chain" in xsmm.last_kernel()
    del probe
    # one launch per tile invoke (no regrouping): keep the rotation short, it is ~2 ms per forward pass
    wl = bench.MlpWorkload(256, to_dev(x), [to_dev(W) for W in Ws], [to_dev(b) for b in bs], tiles=tiles, vnni=vnni,
                           max_sets=None if fused else 4)
    wl.rotations(2)
    torch.cuda.synchronize()
    n = 10 if fused else 2
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = xsmm.launch_count()
    e0.record(stream)
    wl.rotations(n)
    e1.record(stream)
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3 / (n * wl.num_sets)
    launches = (xsmm.launch_count() - l0) / (n * wl.num_sets)
    if os.environ.get("TPP_XSMM_TC_TRACE") == "4":
        xsmm.LIB.xsmm_cuda_debug_dump_trace()
    s = wl.num_sets - 1
    stream_kernel = xsmm.last_kernel()
    rel = bench.rel_err(wl.output(s).cpu().numpy().view(np.uint16), np.roll(bench.oracle_forward(x, Ws, bs), s, 0))
    flops = wl.cfg.flops()
    bn, bk, bc = tiles
    invokes = 3 * (256 // bn) * (1024 // bk)
    # the same stream as tpp-run's own loop runs it: ONE forward pass re-run on ONE set of buffers (bench.py: latency)
    lone = None
    if fused:
        from tpp_mlir_b200 import harness

        hot = harness.NativeMlpLoop(wl.cfg, wl.replay.handles, wl.sets[:1])
        lone = {}
        for label, run in (("unrolled16", lambda k: hot.run_graph_unrolled(k, 16)), ("one_forward_per_graph", hot.run_graph)):
            run(16)              # both graphs (16 forwards / 1 forward) are captured here, outside the timed region
            for _ in range(10):
                run(1)
            torch.cuda.synchronize()
            t0 = xsmm.perf_start_timer()
            run(1024)
            dt = xsmm.perf_stop_timer(t0) / 1024
            lone[label] = {"ms_per_forward": dt * 1e3, "gflops": flops / dt / 1e9, "kernel": xsmm.last_kernel()}
        lone["rel_err_vs_oracle"] = bench.rel_err(wl.output(0).cpu().numpy().view(np.uint16), bench.oracle_forward(x, Ws, bs))
    return {"config": f"{'reference default stream: ' if tiles == (32, 32, 32) and vnni else ''}MLP 3x1024^2 bf16 batch 256, "
                      f"--tiles={bn},{bk},{bc}{' --vnni=2' if vnni else ''} ({invokes} invokes of {bn}x{bk}x{bc} x batch "
                      f"{1024 // bc} per forward pass, block-packed{', VNNI-2 weights' if vnni else ''})",
            "kernel": stream_kernel, "seconds_per_forward": t, "gflops": flops / t / 1e9,
            "kernel_launches_per_forward": launches, "operand_sets": wl.num_sets, "rel_err_vs_oracle": rel,
            "lone_forward": lone,
            "roofline": {"bound": "hbm", "achieved": 7346176 / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": 7346176 / t / 1e9 / pk["hbm_gbs"], "algorithmic_bytes": 7346176}}


def eltwise(pk, only=None):
    out = []
    m = n = 4096
    nbytes = m * n * 2
    ns = sets_needed(2 * nbytes)
    X = [rnd(m, n) for _ in range(ns)]
    Y = [torch.empty(m * n, dtype=torch.bfloat16, device="cuda") for _ in range(ns)]
    bias = rnd(n)

    def run(name, h, invoke, alg_bytes, iters=60):
        if only and not any(name.startswith(o) for o in only):
            return
        t = timed([lambda i=i: invoke(i) for i in range(ns)], iters, warmup=ns)
        out.append({"config": name, "kernel": xsmm.handle_kernel(h), "seconds": t, "gbytes_per_s": alg_bytes / t / 1e9,
                    "reference_convention_gbs(input bytes only)": nbytes / t / 1e9,
                    "roofline": {"bound": "hbm", "achieved": alg_bytes / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                 "frac": alg_bytes / t / 1e9 / pk["hbm_gbs"], "algorithmic_bytes": alg_bytes}})

    h = xsmm.unary_dispatch(xsmm.UNARY_VNNI2, BF16, m, n, n, n, 0)
    run("cfg4 unary vnni_2 pack 4096x4096 bf16", h,
        lambda i: xsmm.LIB.xsmm_unary_invoke(BF16, h, X[i].data_ptr(), 0, Y[i].data_ptr(), 0), 2 * nbytes)
    hu = xsmm.unary_dispatch(xsmm.UNARY_UNVNNI2_EXT, BF16, m, n, n, n, 0)
    run("cfg4 inverse (VNNI2 -> flat) 4096x4096 bf16", hu,
        lambda i: xsmm.LIB.xsmm_unary_invoke(BF16, hu, X[i].data_ptr(), 0, Y[i].data_ptr(), 0), 2 * nbytes)
    ht = xsmm.unary_dispatch(xsmm.UNARY_TRANSPOSE, BF16, m, n, n, m, 0)
    run("unary transpose 4096x4096 bf16", ht,
        lambda i: xsmm.LIB.xsmm_unary_invoke(BF16, ht, X[i].data_ptr(), 0, Y[i].data_ptr(), 0), 2 * nbytes)
    hr = xsmm.unary_dispatch(xsmm.UNARY_RELU, BF16, m, n, n, n, 0)
    run("unary relu 4096x4096 bf16 (out of place)", hr,
        lambda i: xsmm.LIB.xsmm_unary_invoke(BF16, hr, X[i].data_ptr(), 0, Y[i].data_ptr(), 0), 2 * nbytes)
    hz = xsmm.unary_dispatch(xsmm.UNARY_ZERO, BF16, m, n, n, n, 0)
    run("unary zero 4096x4096 bf16", hz,
        lambda i: xsmm.LIB.xsmm_unary_invoke(BF16, hz, X[i].data_ptr(), 0, Y[i].data_ptr(), 0), nbytes)
    hb = xsmm.binary_dispatch(xsmm.BINARY_ADD, BF16, m, n, n, n, n, xsmm.BINARY_FLAG_BCAST_COL_IN_0)
    run("binary add bias(bcast_col_in0) 4096x4096 bf16", hb,
        lambda i: xsmm.LIB.xsmm_binary_invoke(BF16, hb, bias.data_ptr(), 0, X[i].data_ptr(), 0, Y[i].data_ptr(), 0),
        2 * nbytes)
    return out


def pack(pk):
    """tensor.pack / unpack lowered to one unary TPP per 32x32 tile (the reference's
    benchmarks/mlir/fp32-{pack,unpack}-gemm-operand-*.mlir and the bf16 MLP operands), issued (a) as the reference
    would - one launch per tile - and (b) as a captured graph, where the runtime batches the run into one kernel."""
    import numpy as np  # noqa: F401

    out = []
    F32 = xsmm.F32
    cases = [("fp32 pack operand A 512x1024 (32x32 tiles)", F32, 512, 1024, (0, 1), False),
             ("fp32 pack operand B 1024x512 (outer_dims_perm [1,0])", F32, 1024, 512, (1, 0), False),
             ("fp32 unpack 512x512", F32, 512, 512, (0, 1), True),
             ("bf16 pack 4096x4096 (32x32 tiles, 16384 tiles)", BF16, 4096, 4096, (0, 1), False)]
    stream = torch.cuda.current_stream()
    xsmm.set_stream(stream.cuda_stream)
    for name, dtype, m, n, perm, unpack in cases:
        es = 4 if dtype == F32 else 2
        tdt = torch.float32 if dtype == F32 else torch.int16
        nbytes = m * n * es
        ns = sets_needed(2 * nbytes)
        A = [torch.ones(m * n, dtype=tdt, device="cuda") for _ in range(ns)]
        B = [torch.zeros(m * n, dtype=tdt, device="cuda") for _ in range(ns)]
        rp = harness.PackReplay(dtype, m, n, 32, 32, perm, unpack=unpack)
        graphs = []
        for a, b in zip(A, B):
            with xsmm.graph_capture() as g:
                rp.run(*((b, a) if unpack else (a, b)))
            graphs.append(g)
        kernel = xsmm.last_kernel()
        t_batched = timed([g.launch for g in graphs], 20 * ns, warmup=ns)
        direct_iters = 2 if rp.num_tiles > 4096 else 5
        t_direct = timed([lambda a=a, b=b: rp.run(*((b, a) if unpack else (a, b))) for a, b in zip(A, B)], direct_iters,
                         warmup=1)
        for g in graphs:
            g.destroy()
        out.append({"config": f"8f-3 {name}", "kernel": kernel, "tiles": rp.num_tiles, "seconds_batched": t_batched,
                    "seconds_one_launch_per_tile": t_direct, "speedup": t_direct / t_batched,
                    "roofline": {"bound": "hbm", "achieved": 2 * nbytes / t_batched / 1e9, "peak": pk["hbm_gbs"],
                                 "unit": "GB/s", "frac": 2 * nbytes / t_batched / 1e9 / pk["hbm_gbs"],
                                 "algorithmic_bytes": 2 * nbytes}})
    return out


def pack_4096(pk):
    """SURVEY 8f-3 at BASELINE size: tensor.pack of a 4096 x 4096 bf16 matrix into 32 x 32 tiles as the reference lowers it
    (16384 unary identity invokes), captured: one launch of the TMA-to-TMA grid copy."""
    m = n = 4096
    nbytes = m * n * 2
    ns = sets_needed(2 * nbytes)
    A = [torch.randint(0, 30000, (m * n,), dtype=torch.int16, device="cuda") for _ in range(ns)]
    B = [torch.zeros(m * n, dtype=torch.int16, device="cuda") for _ in range(ns)]
    rp = harness.PackReplay(BF16, m, n, 32, 32, (0, 1))
    stream = torch.cuda.current_stream()
    xsmm.set_stream(stream.cuda_stream)
    graphs = []
    for a, b in zip(A, B):
        with xsmm.graph_capture() as g:
            rp.run(a, b)
        graphs.append(g)
    kernel = xsmm.last_kernel()
    t = timed([g.launch for g in graphs], 20 * ns, warmup=ns)
    ok = bool(torch.equal(B[0].reshape(m // 32, n // 32, 32, 32), A[0].reshape(m // 32, 32, n // 32, 32).permute(0, 2, 1, 3)))
    for g in graphs:
        g.destroy()
    return {"config": "8f-3 bf16 pack 4096x4096, 16384 tile invokes of 32x32 captured", "kernel": kernel, "tiles": rp.num_tiles,
            "seconds": t, "gbytes_per_s": 2 * nbytes / t / 1e9, "bit_exact": ok,
            "roofline": {"bound": "hbm", "achieved": 2 * nbytes / t / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": 2 * nbytes / t / 1e9 / pk["hbm_gbs"], "algorithmic_bytes": 2 * nbytes}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only", default="")
    ap.add_argument("--tiles", default="32,32,32")
    ap.add_argument("--vnni", type=int, default=1)
    ap.add_argument("--bench-extras", action="store_true",
                    help="bench.py's side measurements (configs[1], configs[3], the reference default stream, the 4096^2 "
                         "tile-wise pack) as ONE JSON dict on the last line of stdout")
    args = ap.parse_args()
    pk = peaks()
    torch.cuda.set_device(0)
    if args.bench_extras:
        out = {}
        for name, fn in (("cfg2_brgemm_1024x16", cfg2), ("cfg4_vnni2_pack_4096", cfg4),
                         ("reference_default_stream", reference_stream), ("pack_4096_tilewise", pack_4096)):
            try:
                out[name] = fn(pk)
            except Exception as e:   # a side measurement must not take the others down with it
                out[name] = {"error": repr(e)}
        print(json.dumps(out))
        return
    rows = []
    want = set(args.only.split(",")) if args.only else None

    def on(name):
        return want is None or name in want

    if on("cfg2"):
        rows.append(cfg2(pk))
    if on("cfg3b"):
        rows.append(cfg3b(pk))
    if on("cfg4"):
        rows.extend(eltwise(pk))
    if on("refstream"):
        rows.append(reference_stream(pk, tuple(int(v) for v in args.tiles.split(",")), bool(args.vnni)))
    if on("pack"):
        rows.extend(pack(pk))
    if on("cfg5"):
        rows.append(mlp(pk, 2048, "cfg5 MLP 3x1024^2 batch 2048 on ONE GPU (tiles 2048,1024,1024)", (2048, 1024, 1024)))
        rows.append(mlp(pk, 256, "cfg3 MLP 3x1024^2 batch 256 (graph replay, same as bench.py)", (256, 1024, 1024)))
    for r in rows:
        r["peaks"] = pk["source"]
        print(json.dumps(r))
    if args.out:
        with open(args.out, "w") as f:
            for r in rows:
                f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
