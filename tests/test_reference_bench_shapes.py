"""Every matmul / fully-connected / MLP shape the reference benchmarks (benchmarks/config/{fc,matmul,omp,base}/*.json,
the `mlir-gen` command lines; extracted by tests/golden/make_bench_shapes.py into tests/golden/reference_bench_shapes.json),
replayed as the call stream the reference pipeline lowers it to (block-packed operands, one
[fused_]brgemm invoke per output block, batch-reduce over the input blocks) through the C-ABI under graph capture, and
compared with the oracle. f32, bf16 with VNNI-2 and VNNI-4 weights, with and without bias + ReLU; widths 352 ... 4096,
48-wide tiles, batches 128 / 256 / 1024.
"""
from __future__ import annotations

import json
import os

import numpy as np
import pytest

import oracle

F32, BF16 = 1, 2
HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_bench_shapes.json")) as _f:
    SHAPES = json.load(_f)

ROWS = 32   # rows at either end of the batch the oracle recomputes (rows are independent through the layers)


def _resolved(shape):
    """(dtype, vnni factor, tiles) as the lowered program has them: without --tiles the default pipeline packs with
    32 x 32 x 32 blocks (lib/TPP/Transforms/ToBlockLayoutAndBack.cpp:459-471); a bf16 kernel without --vnni gets its
    VNNI layout from the pipeline, with the factor libxsmm_cpuid_dot_pack_factor reports (2; VNNIUtils.cpp:31-37)."""
    dtype = BF16 if shape["float_type"] == "bf16" else F32
    vnni = (shape["vnni"] or 2) if dtype == BF16 else 0
    tiles = tuple(shape["tiles"] or (32, 32, 32))
    return dtype, vnni, tiles


def _id(entry):
    s = entry["shape"]
    return "{}{}-b{}-{}-t{}-{}".format(s["float_type"], f"v{s['vnni']}" if s["vnni"] else "", s["batch"],
                                      "x".join(map(str, s["layers"])), "x".join(map(str, s["tiles"] or ["default"])),
                                      "fc" if s["bias"] else "mm")


def test_fixture_covers_the_reference_benchmark_configs():
    """118 distinct mlir-gen shapes; each one is a legal block-packed MLP of the harness (sizes divide by the tiles) and
    lists the config files that hold it. FLOPs follow mlir-gen's count (tools/mlir-gen/MLIRGen.cpp:313-334), pinned by
    the reference on --batch=128 --layers=2304,768 with bias + ReLU: 453181440 (test/Integration/mlir-gen-named.mlir:5)."""
    from tpp_mlir_b200 import harness

    assert len(SHAPES) == 118
    kinds = set()
    for e in SHAPES:
        s = e["shape"]
        dtype, vnni, tiles = _resolved(s)
        cfg = harness.MlpConfig(batch=s["batch"], layers=tuple(s["layers"]), tiles=tiles, dtype=dtype, vnni=bool(vnni),
                                bias=s["bias"], relu=s["relu"])
        assert cfg.num_layers in (1, 3) and e["sources"]
        if vnni:
            assert tiles[2] % vnni == 0
        kinds.add((dtype, vnni, s["bias"]))
    assert kinds == {(F32, 0, False), (F32, 0, True), (BF16, 2, False), (BF16, 2, True), (BF16, 4, False), (BF16, 4, True)}
    ref = harness.MlpConfig(batch=128, layers=(2304, 768), tiles=(64, 48, 64), dtype=F32, bias=True, relu=True)
    assert ref.flops() == 453181440


@pytest.mark.gpu
@pytest.mark.parametrize("entry", SHAPES, ids=_id)
def test_reference_benchmark_shape_matches_the_oracle(entry, monkeypatch):
    import torch

    from tpp_mlir_b200 import harness, xsmm

    s = entry["shape"]
    dtype, vnni, tiles = _resolved(s)
    bn, bk, bc = tiles
    batch, layers = s["batch"], tuple(s["layers"])
    if vnni == 4:
        monkeypatch.setenv("TPP_XSMM_VNNI", "4")   # what libxsmm_cpuid_dot_pack_factor answers under mlir-gen --vnni=4
    cfg = harness.MlpConfig(batch=batch, layers=layers, tiles=tiles, dtype=dtype, vnni=bool(vnni), bias=s["bias"],
                            relu=s["relu"])
    gen = oracle.TensorInit("normal", dtype, 123)
    Ws = [gen.fill(c, k) for c, k in zip(layers[:-1], layers[1:])]
    bs = [gen.fill(k) for k in layers[1:]]
    x = gen.fill(batch, layers[0])
    tdt = torch.float32 if dtype == F32 else torch.int16

    def t(a):
        return torch.from_numpy(a if dtype == F32 else a.view(np.int16))

    wp = [harness.pack_weight(t(W), bk, bc) for W in Ws]
    if vnni:
        wp = [harness.vnni_pack_weight(w, vnni) for w in wp]
    acts = [harness.pack_activation(t(x), bn, bc).cuda()] + [torch.zeros(batch * k, dtype=tdt).cuda() for k in layers[1:]]
    r = harness.MlpReplay(cfg, [w.cuda() for w in wp], [t(b).cuda() for b in bs], acts)
    n0 = xsmm.launch_count()
    with xsmm.graph_capture() as g:
        r.forward()
    kernel = xsmm.last_kernel()
    g.launch()
    xsmm.sync()
    launches = xsmm.launch_count() - n0
    g.destroy()
    rows = np.r_[0:ROWS, batch - ROWS:batch]
    oracle.set_acc_mode(1 if dtype == F32 else 0)   # f32: f64 accumulation, the reference's summation order is unspecified
    try:
        ref = np.ascontiguousarray(x[rows])
        for W, b in zip(Ws, bs):
            c, k = W.shape
            y = np.zeros((len(rows), k), np.float32 if dtype == F32 else np.uint16)
            if s["bias"] or s["relu"]:
                oracle.fused_brgemm(dtype, len(rows), k, c, c, k, k, 0, 0, 4, 0, 5 if s["relu"] else 0,
                                    4 if s["bias"] else 0, 1 if s["bias"] else 0, ref, W, y, b, 1)
            else:
                oracle.brgemm(dtype, len(rows), k, c, c, k, k, 0, 0, 4, ref, W, y, 1)
            ref = y
    finally:
        oracle.set_acc_mode(0)
    got = harness.unpack_activation(acts[-1].reshape(batch // bn, layers[-1] // bk, bn, bk)).cpu().numpy()[rows]
    if dtype == F32:
        g64, w64, rtol = got.astype(np.float64), ref.astype(np.float64), 1e-5
    else:
        g64 = oracle.bf16_to_f32(np.ascontiguousarray(got).view(np.uint16)).astype(np.float64)
        w64, rtol = oracle.bf16_to_f32(ref).astype(np.float64), 1e-2
    stats = os.environ.get("TPP_TEST_SHAPE_STATS")
    if stats:
        err = float(np.abs(g64 - w64).max() / max(np.abs(w64).max(), 1e-30))
        with open(stats, "a") as f:
            f.write(json.dumps({"id": _id(entry), "kernel": kernel, "launches": launches,
                                "invokes": r.invokes_per_forward, "max_err_over_max": err}) + "\n")
    # never one launch per tile invoke: a layer is one launch (plus, for VNNI-4 weights, one flat copy in front)
    assert launches <= 2 * cfg.num_layers, (kernel, launches, r.invokes_per_forward)
    np.testing.assert_allclose(g64, w64, rtol=rtol, atol=rtol * 0.5 * max(np.abs(w64).max(), 1e-30), err_msg=kernel)
