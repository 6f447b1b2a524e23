"""CPU-only tests of the host-side replay logic (harness) - offsets, flops, layouts."""
import numpy as np
import pytest
import torch

from tpp_mlir_b200 import harness


def test_flops_match_reference_convention():
    # benchmarks/mlir/pytorch/torch-dynamo-mlp-bf16-3x1024.mlir:4 and ...gemm-bf16-3x1024.mlir:4
    cfg = harness.MlpConfig(batch=256, layers=(1024,) * 4, tiles=(256, 1024, 1024))
    assert cfg.flops() == 1_612_185_600
    assert cfg.matmul_flops() == 1_610_612_736
    cfg8 = harness.MlpConfig(batch=2048, layers=(1024,) * 4, tiles=(256, 1024, 1024))
    assert cfg8.flops() == 8 * 1_612_185_600


def test_layouts_and_flops_match_the_reference_mlir_gen_tests():
    """The reference pins mlir-gen's operand types and BENCH_TOTAL_FLOPS in its own tests; the harness must build the
    same block-packed shapes and count the same FLOPs."""
    # test/Integration/mlir-gen-matmul.mlir:1-11, test/BF16/Integration/mlir-gen-matmul-bf16.mlir:1-40 and the -fc twins
    # (mlir-gen-fc.mlir:1-11, mlir-gen-fc-bf16.mlir:1-56): --batch=128 --layers=2304,768 --tiles=64,48,64
    bn, bk, bc = 64, 48, 64
    x, w = torch.zeros(128, 2304), torch.zeros(2304, 768)
    xp, wp = harness.pack_activation(x, bn, bc), harness.pack_weight(w, bk, bc)
    assert tuple(xp.shape) == (2, 36, 64, 64)                                  # %arg0: tensor<2x36x64x64x..>
    assert tuple(wp.shape) == (16, 36, 64, 48)                                 # %arg1: tensor<16x36x64x48x..>
    assert tuple(harness.vnni_pack_weight(wp, 2).shape) == (16, 36, 32, 48, 2)   # DP2: tensor<16x36x32x48x2xbf16>
    assert tuple(harness.vnni_pack_weight(wp, 4).shape) == (16, 36, 16, 48, 4)   # DP4: tensor<16x36x16x48x4xbf16>
    out = torch.zeros(2, 16, 64, 48)                                           # result: tensor<2x16x64x48x..>, bias 16x48
    assert tuple(harness.unpack_activation(out).shape) == (128, 768)
    mm = harness.MlpConfig(batch=128, layers=(2304, 768), tiles=(bn, bk, bc), bias=False, relu=False)
    fc = harness.MlpConfig(batch=128, layers=(2304, 768), tiles=(bn, bk, bc))
    assert mm.flops() == 452984832 and fc.flops() == 453181440
    # test/Integration/mlir-gen-flops.mlir:2-50
    def flops(batch, layers, fused):
        t = (1, 1, 1)
        return harness.MlpConfig(batch=batch, layers=layers, tiles=t, dtype=1, bias=fused, relu=fused).flops()

    assert flops(1, (1, 1), False) == 2 and flops(1, (1, 1), True) == 4                       # MATMUL-UNIT, FC-UNIT / MLP-UNIT
    assert flops(8, (4, 16), False) == 1024 and flops(8, (4, 16), True) == 1280               # MATMUL-SMALL, FC-SMALL
    assert flops(8, (4, 8, 16), True) == 2944                                                 # MLP-SMALL
    assert flops(128, (1024, 4096), False) == 1073741824 and flops(128, (1024, 4096), True) == 1074790400   # *-LARGE
    assert flops(128, (1024, 1024, 1024), True) == 537395200                                  # MLP-LARGE


def test_config_validation():
    with pytest.raises(ValueError):
        harness.MlpConfig(batch=100, tiles=(32, 32, 32))
    with pytest.raises(ValueError):
        harness.MlpConfig(batch=256, layers=(1024, 1024, 1024), tiles=(32, 64, 32))


def test_pack_unpack_roundtrip_and_block_math():
    torch.manual_seed(0)
    mb, c, k, bn, bk, bc = 64, 96, 128, 32, 32, 32
    x, w = torch.rand(mb, c), torch.rand(c, k)
    xp, wp = harness.pack_activation(x, bn, bc), harness.pack_weight(w, bk, bc)
    assert xp.shape == (mb // bn, c // bc, bn, bc) and wp.shape == (k // bk, c // bc, bc, bk)
    # blocked BRGEMM == flat matmul: out[iN][iK] = sum_iC xp[iN][iC] @ wp[iK][iC]
    out = torch.einsum("ncab,kcbd->nkad", xp, wp)
    torch.testing.assert_close(harness.unpack_activation(out), x @ w, rtol=1e-5, atol=1e-5)
    wv = harness.vnni_pack_weight(wp)
    assert wv.shape == (k // bk, c // bc, bc // 2, bk, 2)
    assert wv[1, 2, 3, 4, 1] == wp[1, 2, 7, 4]


def test_replay_offsets_follow_appendix_b(monkeypatch):
    """The invoke stream must be exactly the one the lowered IR issues."""
    from tpp_mlir_b200 import xsmm

    calls = []
    monkeypatch.setattr(xsmm, "fused_brgemm_dispatch", lambda *a: calls.append(("dispatch",) + a) or 77)
    monkeypatch.setattr(xsmm, "intel_amx_tile_config_dispatch", lambda *a: 78)
    monkeypatch.setattr(xsmm, "fused_brgemm_invoke", lambda *a: calls.append(("invoke",) + a))
    cfg = harness.MlpConfig(batch=64, layers=(64, 128), tiles=(32, 32, 32))
    r = harness.MlpReplay(cfg, weights=["W"], biases=["b"], acts=["x", "y"])
    r.forward()
    d = calls[0]
    # dtype, m=bn, n=bk, k=bc, lda=bc, ldb=bk, ldc=bk, stride_a=bn*bc, stride_b=bc*bk, beta_0|64|128, 0, relu, col_in0, add
    assert d[1:] == (2, 32, 32, 32, 32, 32, 32, 1024, 1024, 4 | 64 | 128, 0, 5, 4, 1)
    inv = [c for c in calls if c[0] == "invoke"]
    assert len(inv) == (64 // 32) * (128 // 32) == r.invokes_per_forward
    # (iN=1, iK=2): offA = iN*(C/bc)*bn*bc, offB = iK*(C/bc)*bc*bk, offC = (iN*(K/bk)+iK)*bn*bk, offD = iK*bk
    c = inv[1 * 4 + 2]
    assert c[1:] == (2, 77, "x", 1 * 2 * 1024, "W", 2 * 2 * 1024, "y", (1 * 4 + 2) * 1024, "b", 64, 2)


def test_tensor_pack_oracle_matches_the_tile_loop():
    """oracle.tensor_pack / tensor_unpack (numpy reshapes) against the explicit per-tile loop the reference's lowering
    executes (one 2-D copy per tile, LowerPacksAndUnpacks.cpp:143-250), for both outer-dims permutations."""
    import numpy as np

    import oracle

    rng = np.random.default_rng(3)
    for (m, n, bm, bn) in ((64, 96, 32, 32), (24, 40, 12, 8), (512, 1024, 32, 32)):
        x = rng.integers(0, 1 << 16, size=(m, n), dtype=np.uint16)
        for perm in ((0, 1), (1, 0)):
            p = oracle.tensor_pack(x, bm, bn, perm)
            mb, nb = m // bm, n // bn
            assert p.shape == ((nb, mb, bm, bn) if perm == (1, 0) else (mb, nb, bm, bn))
            for i in range(mb):
                for j in range(nb):
                    tile = p[j, i] if perm == (1, 0) else p[i, j]
                    assert np.array_equal(tile, x[i * bm:(i + 1) * bm, j * bn:(j + 1) * bn])
            assert np.array_equal(oracle.tensor_unpack(p, perm), x)


def test_tile_hazard_test_is_exact_for_equal_pitches():
    """The rectangle-overlap test that guards the batching of captured tile moves (runtime.cu rects_overlap, exported
    as a debug hook) against a brute-force byte-set comparison: random rectangles of one pitch, including rows that
    wrap around the pitch; with different pitches it may only err on the safe side."""
    import ctypes

    import numpy as np

    from tpp_mlir_b200 import _build

    lib = ctypes.CDLL(_build.build())
    fn = lib.xsmm_cuda_debug_rects_overlap
    fn.restype = ctypes.c_int64
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64] * 2
    base = 1 << 20

    def bytes_of(off, rows, width, ld):
        return {off + r * ld + c for r in range(rows) for c in range(width)}

    rng = np.random.default_rng(11)
    seen_true = seen_false = 0
    for _ in range(3000):
        ld = int(rng.integers(8, 40))
        ra, wa, rb, wb = (int(rng.integers(1, 6)), int(rng.integers(1, ld + 1)), int(rng.integers(1, 6)),
                          int(rng.integers(1, ld + 1)))
        oa, ob = int(rng.integers(0, 6 * ld)), int(rng.integers(0, 6 * ld))
        want = bool(bytes_of(oa, ra, wa, ld) & bytes_of(ob, rb, wb, ld))
        got = bool(fn(base + oa, ra, wa, ld, base + ob, rb, wb, ld))
        assert got == want, (ld, oa, ra, wa, ob, rb, wb)
        seen_true += want
        seen_false += not want
    assert seen_true > 300 and seen_false > 300
    for _ in range(1000):   # different pitches: never a false negative
        lda, ldb = int(rng.integers(8, 40)), int(rng.integers(8, 40))
        ra, wa, rb, wb = (int(rng.integers(1, 6)), int(rng.integers(1, lda + 1)), int(rng.integers(1, 6)),
                          int(rng.integers(1, ldb + 1)))
        oa, ob = int(rng.integers(0, 200)), int(rng.integers(0, 200))
        if bytes_of(oa, ra, wa, lda) & bytes_of(ob, rb, wb, ldb):
            assert fn(base + oa, ra, wa, lda, base + ob, rb, wb, ldb) == 1
    # the pack / unpack pattern: 32 x 32 f32 tiles of a 1024-wide matrix are pairwise disjoint
    ld, w = 1024 * 4, 32 * 4
    for (i, j) in ((0, 1), (1, 0), (1, 1), (0, 31)):
        assert fn(base, 32, w, ld, base + i * 32 * ld + j * w, 32, w, ld) == 0
    assert fn(base, 32, w, ld, base + 31 * ld + w - 1, 32, w, ld) == 1


def test_bench_keeps_library_chatter_off_stdout():
    """bench.py's stdout contract is ONE JSON line: whatever libraries write to fd 1 while it runs (NCCL prints its
    version banner there) is diverted to stderr until the line itself is printed."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = "\n".join([
        "import os, sys",
        "sys.path.insert(0, %r)" % root,
        "import bench",
        "fd = bench.capture_stdout()",
        "print('python noise')",
        "os.write(1, b'native noise' + bytes([10]))",
        "bench.restore_stdout(fd)",
        "print('{}')",
    ])
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == "{}\n"
    assert "python noise" in r.stderr and "native noise" in r.stderr


def test_bench_reference_arm_prints_the_contract_line_on_rank_0_only():
    """`bench.py --impl reference` (the driver's CPU arm): rank 0 prints ONE JSON line with the GPU arm's metric, unit and
    workload string, its own cpu_baseline / e2e objects and zero device copies; other ranks print nothing and exit 0."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench

    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "3"]
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout == "", (r.stdout, r.stderr)
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="2")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["n_gpus"] == 2 and d["steps"] == 2 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["config"]["workload"] == bench.workload_text(bench.BATCH_PER_GPU) and d["config"]["global_batch"] == 512
    assert d["value"] > 0 and d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["cores"] >= 1 and "rows" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


# ---- folding of captured tile invokes into layers (runtime.cu: fold_grid), no GPU needed --------------------------------
def _fold(m, n, k, lda, ldb, ldc, sa, sb, flags, batch, invokes, with_bias=True):
    import ctypes

    from tpp_mlir_b200 import xsmm

    num = len(invokes)
    arr = lambda col: (ctypes.c_int64 * num)(*[inv[col] for inv in invokes])   # noqa: E731
    out = (ctypes.c_int64 * 8)()
    xsmm.LIB.xsmm_cuda_debug_fold_grid(m, n, k, lda, ldb, ldc, sa, sb, flags, batch, num, arr(0), arr(1), arr(2),
                                       arr(3) if with_bias else None, out)
    return dict(zip(("grid_n", "grid_k", "a_step", "b_step", "c_step_n", "c_step_k", "d_step", "folded"), out))


def _appendix_b_invokes(mb, c, kk, bn, bk, bc, order="nk"):
    """operand offsets of one layer of the block-packed MLP, (iN outer, iK inner) or reversed"""
    nb_c, nb_k = c // bc, kk // bk
    pairs = [(i, j) for i in range(mb // bn) for j in range(nb_k)] if order == "nk" else \
            [(i, j) for j in range(nb_k) for i in range(mb // bn)]
    return [(i * nb_c * bn * bc, j * nb_c * bc * bk, (i * nb_k + j) * bn * bk, j * bk) for i, j in pairs]


def test_fold_reference_default_layer():
    """benchmarks/config/omp/mlir-bf16.json:37 (--tiles=32,32,32): 256 invokes per layer fold into an 8 x 32 grid"""
    inv = _appendix_b_invokes(256, 1024, 1024, 32, 32, 32)
    r = _fold(32, 32, 32, 32, 32, 32, 1024, 1024, 4 | 2048, 32, inv)
    assert r == {"grid_n": 8, "grid_k": 32, "a_step": 32 * 1024, "b_step": 32 * 1024, "c_step_n": 32 * 1024, "c_step_k": 1024,
                 "d_step": 32, "folded": 256}


def test_fold_stops_at_the_layer_boundary_and_handles_both_loop_orders():
    layer0 = _appendix_b_invokes(256, 512, 512, 64, 64, 64)
    # the next layer reads what this one wrote (different A / B / C bases): must not be folded into the same grid
    layer1 = [(a + 10_000_000, b + 20_000_000, c + 30_000_000, d + 5_000) for a, b, c, d in layer0]
    r = _fold(64, 64, 64, 64, 64, 64, 4096, 4096, 4, 8, layer0 + layer1)
    assert (r["grid_n"], r["grid_k"], r["folded"]) == (4, 8, 32)
    r = _fold(64, 64, 64, 64, 64, 64, 4096, 4096, 4, 8, _appendix_b_invokes(256, 512, 512, 64, 64, 64, order="kn"))
    assert (r["grid_n"], r["grid_k"], r["folded"]) == (4, 8, 32)
    assert r["a_step"] == 8 * 4096 and r["b_step"] == 8 * 4096 and r["c_step_k"] == 4096 and r["c_step_n"] == 8 * 4096


def test_fold_refuses_irregular_or_hazardous_walks():
    inv = _appendix_b_invokes(256, 512, 512, 64, 64, 64)
    shuffled = [inv[i] for i in (2, 0, 3, 1)] + inv[4:]
    assert _fold(64, 64, 64, 64, 64, 64, 4096, 4096, 4, 8, shuffled)["folded"] == 1
    # output tiles that overlap each other (c step smaller than a tile): not a layer
    bad = [(a, b, c // 2, d) for a, b, c, d in inv]
    assert _fold(64, 64, 64, 64, 64, 64, 4096, 4096, 4, 8, bad)["folded"] == 1
    # a single invoke, and two invokes that share neither A nor B
    assert _fold(256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4, 1, [(0, 0, 0, 0)])["folded"] == 1
    assert _fold(64, 64, 64, 64, 64, 64, 4096, 4096, 4, 8, [(0, 0, 0, 0), (4096 * 8, 4096 * 8, 4096, 64)])["folded"] == 1
    # flat layouts tiled along n only (one row block): interleaved output tiles of one matrix are fine
    flat = [(0, j * 64, j * 64, j * 64) for j in range(8)]
    r = _fold(256, 64, 512, 512, 512, 512, 0, 0, 4, 1, flat)
    assert (r["grid_n"], r["grid_k"], r["c_step_k"], r["folded"]) == (1, 8, 64, 8)


def test_fold_f32_tile_invokes():
    """fp32 tile invokes (the reference's fp32 MLP configs) fold into the same 8 x 32 grid: steps are in 4-byte elements"""
    inv = _appendix_b_invokes(256, 1024, 1024, 32, 32, 32)
    r = _fold(32, 32, 32, 32, 32, 32, 1024, 1024, 4 | (1 << 40), 32, inv)
    assert r == {"grid_n": 8, "grid_k": 32, "a_step": 32 * 1024, "b_step": 32 * 1024, "c_step_n": 32 * 1024, "c_step_k": 1024,
                 "d_step": 32, "folded": 256}


def _tile_grid(moves):
    import ctypes

    from tpp_mlir_b200 import xsmm

    num = len(moves)
    a = (ctypes.c_int64 * num)(*[m[0] for m in moves])
    b = (ctypes.c_int64 * num)(*[m[1] for m in moves])
    out = (ctypes.c_int64 * 6)()
    ok = xsmm.LIB.xsmm_cuda_debug_tile_grid(num, a, b, out)
    return bool(ok), dict(zip(("J", "I", "in_inner", "in_outer", "out_inner", "out_outer"), out))


def test_tile_runs_of_a_lowered_pack_are_regular_grids():
    """tensor.pack of a 256 x 1024 bf16 matrix into 32 x 32 tiles, tile by tile (harness.PackReplay order): the run is a
    regular 8 x 32 grid for the TMA-to-TMA copy - also with outer_dims_perm = [1, 0] and for the unpack direction; a
    shuffled run is not."""
    M, N, bm, bn, es = 256, 1024, 32, 32, 2
    mb, nb = M // bm, N // bn
    flat = lambda i, j: (i * bm * N + j * bn) * es            # noqa: E731
    packed = lambda i, j: (i * nb + j) * bm * bn * es         # noqa: E731
    packed_t = lambda i, j: (j * mb + i) * bm * bn * es       # noqa: E731
    order = [(i, j) for i in range(mb) for j in range(nb)]
    ok, g = _tile_grid([(flat(i, j), packed(i, j)) for i, j in order])
    assert ok and g == {"J": 32, "I": 8, "in_inner": 64, "in_outer": 32 * 2048, "out_inner": 2048, "out_outer": 32 * 2048}
    ok, g = _tile_grid([(flat(i, j), packed_t(i, j)) for i, j in order])
    assert ok and (g["J"], g["I"], g["out_inner"], g["out_outer"]) == (32, 8, mb * 2048, 2048)
    ok, g = _tile_grid([(packed(i, j), flat(i, j)) for i, j in order])
    assert ok and (g["J"], g["I"], g["in_inner"], g["out_inner"]) == (32, 8, 2048, 64)
    shuffled = [order[k] for k in (1, 0, 2, 3)] + order[4:]
    assert not _tile_grid([(flat(i, j), packed(i, j)) for i, j in shuffled])[0]
    # one row of tiles only: J = the whole run, I = 1
    ok, g = _tile_grid([(flat(0, j), packed(0, j)) for j in range(nb)])
    assert ok and (g["J"], g["I"]) == (32, 1)
