"""The reference's known-answer tests, replayed op by op on a backend.

Each case issues the xsmm calls its reference test contains (dispatch arguments
are copied from the cited .mlir, inputs/expected numbers come from
tests/golden/reference_vectors.json, which tests/golden/make_golden.py extracted
from those same files) and returns (result, expected, abs_tol). The same cases
pin the oracle (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_parity_gpu.py, through the C-ABI).
"""
from __future__ import annotations

import json
import os

import numpy as np

import oracle

F32, BF16 = 1, 2
_G = None


def golden():
    global _G
    if _G is None:
        with open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")) as f:
            _G = json.load(f)
    return _G


def const(dtype, shape, value=1.0):
    """tpp-run's default input init without --seed: dense<1.0> (TensorInit.cpp:77-82)."""
    a = np.full(shape, value, dtype=np.float32)
    return a if dtype == F32 else oracle.f32_to_bf16(a)


def to_f32(dtype, a):
    return a.astype(np.float32) if dtype == F32 else oracle.bf16_to_f32(a)


def from_f32(dtype, a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a if dtype == F32 else oracle.f32_to_bf16(a)


# ---- cases -----------------------------------------------------------------------

def case_brgemm_f32_ones(be):
    # test/Integration/xsmm-brgemm.mlir:13: dispatch [32,64,16,16,64,64,512,1024] flags none, batch 2
    A, B, C = const(F32, (2, 32, 16)), const(F32, (2, 16, 64)), const(F32, (64 * 32,))
    be.brgemm(F32, 32, 64, 16, 16, 64, 64, 512, 1024, 0, A, 0, B, 0, C, 0, 2)
    return C, np.array(golden()["brgemm_f32_ones"]["expected"], np.float32), 0.0


def case_brgemm_bf16_vnni(be):
    # test/BF16/Integration/xsmm-brgemm-bf16.mlir:12: [6,6,6,6,6,6,36,36] flags (vnni_b) -> 2048 at the ABI
    A, B, C = const(BF16, (2, 6, 6)), const(BF16, (2, 3, 6, 2), 3.0), const(BF16, (36,))
    be.brgemm(BF16, 6, 6, 6, 6, 6, 6, 36, 36, 2048, A, 0, B, 0, C, 0, 2)
    return to_f32(BF16, C), np.array(golden()["brgemm_bf16_vnni"]["expected"], np.float32), 0.0


def case_brgemm_bf16_vnni_batch64(be):
    # test/BF16/Integration/xsmm-ternary-bf16.mlir:8: [4,4,4,4,4,4,8,8] (vnni_b), batch 64; 257 rounds to 256
    A, B, C = const(BF16, (64, 4, 4)), const(BF16, (64, 2, 4, 2)), const(BF16, (16,))
    be.brgemm(BF16, 4, 4, 4, 4, 4, 4, 8, 8, 2048, A, 0, B, 0, C, 0, 64)
    return to_f32(BF16, C), np.array(golden()["brgemm_bf16_vnni_batch64"]["expected"], np.float32), 0.0


def case_gemm_bf16_vnni(be):
    # test/BF16/Integration/xsmm-gemm-bf16.mlir:10: gemm [6,6,6,6,6,6] (vnni_b)
    A, B, C = const(BF16, (6, 6)), const(BF16, (3, 6, 2), 3.0), const(BF16, (36,))
    be.gemm(BF16, 6, 6, 6, 6, 6, 6, 2048, A, 0, B, 0, C, 0)
    return to_f32(BF16, C), np.array(golden()["gemm_bf16_vnni"]["expected"], np.float32), 0.0


def _case_fused(be, dtype, gflags, key):
    # test/Integration/xsmm-quarternary.mlir:6-9 / BF16 twin: [4,4,4,4,4,4,8,8][add,relu],
    # binary_flags (bcast_col_in0), batch 16, C accumulated (no beta_0): 16*4 + 1 + 1 = 66
    A = const(dtype, (64, 4, 4))
    B = const(dtype, (64, 4, 4)) if dtype == F32 else const(dtype, (64, 2, 4, 2))
    C, D = const(dtype, (16,)), const(dtype, (4,))
    be.fused_brgemm(dtype, 4, 4, 4, 4, 4, 4, 8, 8, gflags, 0, 5, 4, 1, A, 0, B, 0, C, 0, D, 0, 16)
    g = golden()[key]
    return to_f32(dtype, C), np.full(16, g["expected_fill"], np.float32), g["threshold"]


def case_fused_f32(be):
    return _case_fused(be, F32, 0, "fused_f32")


def case_fused_bf16_vnni(be):
    return _case_fused(be, BF16, 2048, "fused_bf16_vnni")


def case_fused_f32_seed123(be):
    # test/Integration/xsmm-fusion.mlir:51: dispatch (1,4,4,8,8,4,4,32,32,4,0,5,4,1), batch 2, --seed 123:
    # kernel args are filled in order (A 2x4x8, bias 1x4) from one normal generator; B is dense<2.0>
    gen = oracle.TensorInit("normal", F32, 123)
    A, bias = gen.fill(2, 4, 8), gen.fill(1, 4)
    B, C = const(F32, (2, 8, 4), 2.0), np.zeros(16, np.float32)
    be.fused_brgemm(F32, 4, 4, 8, 8, 4, 4, 32, 32, 4, 0, 5, 4, 1, A, 0, B, 0, C, 0, bias, 0, 2)
    return C, np.array(golden()["fused_f32_seed123"]["expected"], np.float32), 6e-6  # 6 printed digits


def case_transpose_f32(be):
    # test/Integration/xsmm-transpose.mlir:6: transpose [4,8,8,4]
    g = golden()["transpose_f32"]
    inp, out = np.array(g["input"], np.float32), np.zeros(32, np.float32)
    be.unary(29, F32, 4, 8, 8, 4, 0, inp, 0, out, 0)
    return out, np.array(g["expected"], np.float32), 1e-6


def case_vnni2_bf16_seed123(be):
    # test/Integration/transpose-bf16.mlir:10-18 (--seed 123): 4x4 bf16 -> [2][4][2] == vnni_2 [4,4,4,4]
    gen = oracle.TensorInit("normal", BF16, 123)
    inp = gen.fill(4, 4)
    out = np.zeros(16, np.uint16)
    be.unary(28, BF16, 4, 4, 4, 4, 0, inp, 0, out, 0)
    return to_f32(BF16, out), np.array(golden()["vnni2_bf16_seed123"]["expected"], np.float32), 1e-6


def case_vnni2_pack_16x16(be):
    # test/BF16/Integration/vnni-packing.mlir:5-8: 16x16 -> 8x16x2, i.e. vnni_2 [16,16,16,16]
    g = golden()["vnni2_pack_16x16"]
    inp, out = from_f32(BF16, np.array(g["input"])), np.zeros(256, np.uint16)
    be.unary(28, BF16, 16, 16, 16, 16, 0, inp, 0, out, 0)
    pref = np.array(g["expected_prefix"], np.float32)
    return to_f32(BF16, out)[:pref.size], pref, 0.0


def case_vnni2_pack_chain(be):
    # test/BF16/Integration/vnni-packing-chain.mlir:7-15: 32x32 -> blocks [2][2][16][16] with
    # outer_dims_perm=[1,0] (a relayout, done here in numpy) -> per block vnni_2 [16,16,16,16]
    g = golden()["vnni2_pack_chain"]
    x = np.array(g["input"], np.float32).reshape(32, 32)
    blocks = x.reshape(2, 16, 2, 16).transpose(2, 0, 1, 3)  # [jb][ib][16][16]
    inp = from_f32(BF16, np.ascontiguousarray(blocks))
    out = np.zeros(1024, np.uint16)
    for blk in range(4):
        be.unary(28, BF16, 16, 16, 16, 16, 0, inp, blk * 256, out, blk * 256)
    return to_f32(BF16, out), np.array(g["expected"], np.float32), 0.0


def _case_unary(be, dtype, kind, key, fill):
    # test/Integration/xsmm-unary.mlir:5 relu [3,3,3,3] in place; xsmm-zero.mlir:8 zero [3,3,3,3] on a 5.0 buffer
    x = const(dtype, (9,), fill)
    be.unary(kind, dtype, 3, 3, 3, 3, 0, x, 0, x, 0)
    return to_f32(dtype, x), np.array(golden()[key]["expected"], np.float32), 0.0


def case_unary_relu_f32(be):
    return _case_unary(be, F32, 5, "unary_relu_f32", 1.0)


def case_unary_relu_bf16(be):
    return _case_unary(be, BF16, 5, "unary_relu_bf16", 1.0)


def case_unary_zero_f32(be):
    return _case_unary(be, F32, 2, "unary_zero_f32", 5.0)


def case_unary_zero_bf16(be):
    return _case_unary(be, BF16, 2, "unary_zero_bf16", 5.0)


def _case_add(be, dtype, key):
    # test/Integration/xsmm-binary.mlir:9: add [3,3,3,3,3]
    a, b, o = const(dtype, (9,)), const(dtype, (9,)), const(dtype, (9,))
    be.binary(1, dtype, 3, 3, 3, 3, 3, 0, a, 0, b, 0, o, 0)
    return to_f32(dtype, o), np.array(golden()[key]["expected"], np.float32), 0.0


def case_binary_add_f32(be):
    return _case_add(be, F32, "binary_add_f32")


def case_binary_add_bf16(be):
    return _case_add(be, BF16, "binary_add_bf16")


def _case_mulsub(be, kind, key):
    # test/Integration/xsmm-mul.mlir:6 / xsmm-sub.mlir:6: [4,8,8,8,8], both operands the same constant
    g = golden()[key]
    x, o = np.array(g["input"], np.float32), np.zeros(32, np.float32)
    be.binary(kind, F32, 4, 8, 8, 8, 8, 0, x, 0, x.copy(), 0, o, 0)
    return o, np.array(g["expected"], np.float32), 1e-5 * 70  # printed with 4-6 significant digits


def case_binary_mul_f32(be):
    return _case_mulsub(be, 2, "binary_mul_f32")


def case_binary_sub_f32(be):
    return _case_mulsub(be, 3, "binary_sub_f32")


def case_binary_div_f32(be):
    # test/Integration/xsmm-div.mlir:6-28: div [4,8,8,8,8] none / (bcast_col_in1) / [4,8,8,1,8] (bcast_row_in1)
    # / [4,8,8,1,8] (bcast_scalar_in1)
    g = golden()["binary_div_f32"]
    lhs = np.array(g["lhs"], np.float32)
    outs = []
    for ldr, flags, key in ((8, 0, "rhs_full"), (8, 8, "rhs_col"), (1, 2, "rhs_row"), (1, 32, "rhs_scalar")):
        rhs, o = np.array(g[key], np.float32), np.zeros(32, np.float32)
        be.binary(4, F32, 4, 8, 8, ldr, 8, flags, lhs, 0, rhs, 0, o, 0)
        outs.append(o)
    return np.concatenate(outs), np.array(g["expected_all"], np.float32), 0.0


def case_strided_brgemm(be):
    # test/Integration/xsmm-strided-brgemm.mlir:34: xsmm_brgemm_dispatch(1,2,2,4,8,16,2,4,64,0):
    # C_exp[i][ii][j][jj] += sum_{k,kk} A_exp[i][ii][k][kk] * B_exp[k][kk][j][jj]; one invoke per (i,j)
    # on a 2x2 tile (ldc=2) that starts at zero, then + D (the linalg.generic add)
    g = golden()["strided_brgemm"]
    A, B, D = (np.array(g[x], np.float32) for x in ("A", "B", "D"))
    res = np.zeros((4, 16), np.float32)
    for i in range(2):
        for j in range(8):
            tile = np.zeros(4, np.float32)
            be.brgemm(F32, 2, 2, 4, 8, 16, 2, 4, 64, 0, A, i * 16, B, j * 2, tile, 0, 2)
            res[i * 2:(i + 1) * 2, j * 2:(j + 1) * 2] = tile.reshape(2, 2)
    res += D.reshape(4, 16)
    return res.reshape(-1), np.array(g["expected"], np.float32), 5e-3  # printed with 2 decimals


def case_strided_gemm1(be):
    # test/Integration/xsmm-strided-brgemm1.mlir:33: xsmm_gemm_dispatch(1,2,2,4,4,16,16,4) (beta_0 folded in):
    # C[b][i][h][j] = sum_k A[b][h][i][k] * B[b][k][h][j]; one gemm per (b,h) writing a strided 2x2 of C(4x16)
    g = golden()["strided_gemm1"]
    A, B = np.array(g["A"], np.float32), np.array(g["B"], np.float32)
    C = np.zeros(64, np.float32)
    for b in range(2):
        for h in range(8):
            be.gemm(F32, 2, 2, 4, 4, 16, 16, 4, A, (b * 8 + h) * 8, B, b * 64 + h * 2, C, b * 32 + h * 2)
    return C, np.array(g["expected"], np.float32), 5e-3 * 10


def case_strided_gemm2(be):
    # test/Integration/xsmm-strided-brgemm2.mlir:35: xsmm_gemm_dispatch(1,2,8,4,8,16,8,4) (beta_0 folded in):
    # C_exp[b][h][i][j] = sum_k A_exp[b][i][h][k] * B_exp[b][k][h][j] with A(4x8) -> (2,2,2,4), B(8x16) -> (2,4,2,8),
    # C(4x16) -> (2,2,2,8): one gemm per (b,h) on A rows at pitch 8, B rows at pitch 16, a 2x8 tile of C at pitch 8
    g = golden()["strided_gemm2"]
    A, B = np.array(g["A"], np.float32), np.array(g["B"], np.float32)
    C = np.zeros(64, np.float32)
    for b in range(2):
        for h in range(2):
            be.gemm(F32, 2, 8, 4, 8, 16, 8, 4, A, b * 16 + h * 4, B, b * 64 + h * 8, C, (b * 2 + h) * 16)
    return C, np.array(g["expected"], np.float32), 5e-3 * 10   # printed with up to 5 significant digits


def case_strided_gemm3(be):
    # test/Integration/xsmm-strided-brgemm3.mlir:34-35: ONE xsmm_unary_dispatch (transpose of the A slice) and
    # xsmm_gemm_dispatch(1,8,2,4,8,2,2,4): C_exp[b][h][j][i] = sum_k A_exp[b][i][h][k] * B_exp[b][j][h][k] with B(16x8) ->
    # (2,8,2,4), C(4x16) -> (2,2,8,2): per (b,h) the 8x4 slice of B is the row-major A operand (pitch 8), the transposed
    # 2x4 slice of A the B operand (4x2), the result an 8x2 tile of C
    g = golden()["strided_gemm3"]
    A, B = np.array(g["A"], np.float32), np.array(g["B"], np.float32)
    C = np.zeros(64, np.float32)
    for b in range(2):
        for h in range(2):
            T = np.zeros(8, np.float32)
            be.unary(29, F32, 2, 4, 8, 2, 0, A, b * 16 + h * 4, T, 0)          # 2x4 (pitch 8) -> 4x2
            be.gemm(F32, 8, 2, 4, 8, 2, 2, 4, B, b * 64 + h * 4, T, 0, C, (b * 2 + h) * 16)
    return C, np.array(g["expected"], np.float32), 5e-3 * 10


def case_mlir_gen_bf16(be):
    # test/BF16/Integration/mlir-gen-bf16.mlir:9-26: mlir-gen --kernel=args --float-type=bf16 --batch=16 --layers=16,16
    # (all-ones init): matmul accumulating into C = 1 prints rows of 17; with --bias --relu rows of 18. VNNI-2 weights, as
    # the bf16 pipeline packs them.
    g = golden()["mlir_gen_bf16"]
    x, W, bias = const(BF16, (16, 16)), const(BF16, (8, 16, 2)), const(BF16, (16,))
    c_mm, c_fc = const(BF16, (16, 16)), const(BF16, (16, 16))
    be.gemm(BF16, 16, 16, 16, 16, 16, 16, 2048, x, 0, W, 0, c_mm, 0)
    be.fused_brgemm(BF16, 16, 16, 16, 16, 16, 16, 0, 0, 2048, 0, 5, 4, 1, x, 0, W, 0, c_fc, 0, bias, 0, 1)
    got = np.concatenate([to_f32(BF16, c_mm)[0], to_f32(BF16, c_fc)[0]])
    return got, np.array(g["matmul_row"] + g["fc_row"], np.float32), 0.0


def case_matmul_64x64x64_f32(be):
    # test/Integration/matmul_64x64x64.mlir:11,16 (cfg1): default 32x32x32 packing -> per output block one
    # xsmm_brgemm_dispatch(1,32,32,32,32,32,32,1024,1024,0) invoke with 2 batches, C initialised to 1 => 65
    A = const(F32, (2, 2, 32, 32))   # [M/32][K/32][32][32]
    B = const(F32, (2, 2, 32, 32))   # [N/32][K/32][32][32]
    C = const(F32, (2, 2, 32, 32))   # [M/32][N/32][32][32]
    for im in range(2):
        for jn in range(2):
            be.brgemm(F32, 32, 32, 32, 32, 32, 32, 1024, 1024, 0, A, im * 2048, B, jn * 2048, C, (im * 2 + jn) * 1024, 2)
    row = golden()["matmul_64x64x64_f32"]["raw_checks"][0]
    return C.reshape(-1), np.full(4096, row["values"][0], np.float32), 0.0


def case_mlp_all_ones_bf16(be):
    # test/BF16/Integration/mlp-all-bf16-tpprun.mlir: batch 128, 256->512->1024->2048->1000, VNNI-2 weights,
    # all inputs/weights/biases 1.0; relu(x*W + b) per layer => 2^38 (expected 2.74878e11, threshold 1.0 bf16)
    g = golden()["mlp_all_ones_bf16"]
    batch, layers = g["batch"], g["layers"]
    x = const(BF16, (batch, layers[0]))
    for c, k in zip(layers[:-1], layers[1:]):
        W = const(BF16, (c // 2, k, 2))
        bias = const(BF16, (k,))
        y = np.zeros((batch, k), np.uint16)
        # fused form of the layer (what CombineXsmmOpPass produces): beta_0 | vnni_b, add bcast_col_in0, relu
        be.fused_brgemm(BF16, batch, k, c, c, k, k, 0, 0, 4 | 2048, 0, 5, 4, 1, x, 0, W, 0, y, 0, bias, 0, 1)
        x = y
    expect = to_f32(BF16, from_f32(BF16, np.full(x.size, g["expected_fill"], np.float32)))
    return to_f32(BF16, x).reshape(-1), expect, g["threshold"]


def case_brgemm_f32_ternary(be):
    # test/Integration/xsmm-ternary.mlir:7: dispatch [3,3,4,4,3,3,12,12] flags none, batch 2, all ones: 1 + 2*4 = 9
    A, B, C = const(F32, (2, 3, 4)), const(F32, (2, 4, 3)), const(F32, (9,))
    be.brgemm(F32, 3, 3, 4, 4, 3, 3, 12, 12, 0, A, 0, B, 0, C, 0, 2)
    return C, np.array(golden()["brgemm_f32_ternary"]["expected"], np.float32), 0.0


def _case_whole_matmul(be, key, m, n, k):
    # M or N no multiple of the default 32 x 32 x 32 packing: the pipeline keeps ONE xsmm_gemm_invoke on the whole
    # matrices (the IR check of the test), all ones, C initialised to 1 => k + 1
    A, B, C = const(F32, (m, k)), const(F32, (k, n)), const(F32, (m * n,))
    be.gemm(F32, m, n, k, k, n, n, 0, A, 0, B, 0, C, 0)
    rows = golden()[key]["raw_checks"]
    assert sum(r["repeat"] for r in rows) == m and all(len(r["values"]) == n for r in rows)
    return C, np.concatenate([np.tile(np.array(r["values"], np.float32), r["repeat"]) for r in rows]), 0.0


def case_matmul_48x64x96_f32(be):
    # test/Integration/matmul_48x64x96.mlir:8-15
    return _case_whole_matmul(be, "matmul_48x64x96_f32", 48, 64, 96)


def case_matmul_64x48x96_f32(be):
    # test/Integration/matmul_64x48x96.mlir:8-15
    return _case_whole_matmul(be, "matmul_64x48x96_f32", 64, 48, 96)


def case_brgemm_f32_non_unit_batch(be):
    # test/Integration/tpp-brgemm-non-unit-batch.mlir:12-16: linalg.batch_reduce_matmul 2x4x8 . 2x8x4 into a zero 4x4
    # -> xsmm_brgemm_dispatch(1, 4, 4, 8, 8, 4, 4, 32, 32, 0), batch 2; non-trivial operands (dense<> literals)
    g = golden()["brgemm_f32_non_unit_batch"]
    A, B, C = np.array(g["A"], np.float32), np.array(g["B"], np.float32), np.zeros(16, np.float32)
    be.brgemm(F32, 4, 4, 8, 8, 4, 4, 32, 32, 0, A, 0, B, 0, C, 0, 2)
    return C, np.array(g["expected"], np.float32), 6e-3   # 6 printed digits of values around 1e3


def case_transpose_f32_seed123(be):
    # test/Integration/transpose-fp32.mlir:7-8 (--seed 123): 3x5 f32 from tpp-run's normal TensorInit -> 5x3; pins the
    # f32 stream of the RNG restatement
    inp = oracle.TensorInit("normal", F32, 123).fill(3, 5)
    out = np.zeros(15, np.float32)
    be.unary(29, F32, 3, 5, 5, 3, 0, inp, 0, out, 0)
    return out, np.array(golden()["transpose_f32_seed123"]["expected"], np.float32), 1e-6


def case_matmul_pbf16(be):
    # test/BF16/Integration/matmul-pbf16.mlir:9-41: A 4x8 ones, B already VNNI-packed [4][4][2] ones, C zeros, plain
    # accumulation: gemm [4,4,8,8,4,4] (vnni_b) => 8
    A, B, C = const(BF16, (4, 8)), const(BF16, (4, 4, 2)), np.zeros(16, np.uint16)
    be.gemm(BF16, 4, 4, 8, 8, 4, 4, 2048, A, 0, B, 0, C, 0)
    return to_f32(BF16, C), np.array(golden()["matmul_pbf16"]["expected"], np.float32), 0.0


def case_mlp_single_layer_blocked_bf16(be):
    # test/BF16/Integration/mlp-single-layer-blocked-bf16.mlir:22-52: per output block a BRGEMM over 64 blocks of 4x4x4
    # (A block stride 16, VNNI-2 B block stride 16) accumulated onto C = 1, then relu in place: 1 + 64*4 = 257 -> 256 in
    # bf16. One (row block, column block) of the test's loop nest; every block prints the same tile
    A, B, C = const(BF16, (64, 4, 4)), const(BF16, (64, 2, 4, 2)), const(BF16, (16,))
    be.brgemm(BF16, 4, 4, 4, 4, 4, 4, 16, 16, 2048, A, 0, B, 0, C, 0, 64)
    be.unary(5, BF16, 4, 4, 4, 4, 0, C, 0, C, 0)
    return to_f32(BF16, C), np.array(golden()["mlp_single_layer_blocked_bf16"]["expected"], np.float32), 0.0


def case_tpp_matmul_f32(be):
    # test/Integration/tpp-matmul.mlir:13-58: a 4x8 . 8x4 contraction with non-trivial dense operands into a zero C
    # -> one xsmm gemm [4,4,8,8,4,4]
    g = golden()["tpp_matmul_f32"]
    A, B, C = np.array(g["A"], np.float32), np.array(g["B"], np.float32), np.zeros(16, np.float32)
    be.gemm(F32, 4, 4, 8, 8, 4, 4, 0, A, 0, B, 0, C, 0)
    return C, np.array(g["expected"], np.float32), 6e-3   # 6 printed digits of values around 5e2


def case_tpp_relu_f32(be):
    # test/Integration/tpp-relu.mlir:14-54: relu of a 9x6 matrix with a block of negative entries, unary relu [9,6,6,6]
    g = golden()["tpp_relu_f32"]
    inp, out = np.array(g["input"], np.float32), np.zeros(54, np.float32)
    be.unary(5, F32, 9, 6, 6, 6, 0, inp, 0, out, 0)
    return out, np.array(g["expected"], np.float32), 1e-6


def case_copy_broadcasts_f32(be):
    # test/Integration/copy.mlir: identity with the three broadcast flags - a 1x6 row into 9x6 (bcast_col = 4: in[0][j]),
    # a 6x1 column into 6x9 (bcast_row = 2: in[i][0], ldi = 1), a 1x1 scalar into 6x9 (bcast_scalar = 8). The third print
    # of the test (23.1) is a linalg.fill, not an xsmm op.
    g = golden()["copy_broadcasts_f32"]
    exp = np.array(g["expected_all"], np.float32)
    assert exp.size == 4 * 54
    o1, o2, o3 = np.zeros(54, np.float32), np.zeros(54, np.float32), np.zeros(54, np.float32)
    be.unary(1, F32, 9, 6, 6, 6, 4, np.array(g["row"], np.float32), 0, o1, 0)
    be.unary(1, F32, 6, 9, 1, 9, 2, np.array(g["col"], np.float32), 0, o2, 0)
    be.unary(1, F32, 6, 9, 1, 9, 8, np.array(g["scalar"], np.float32), 0, o3, 0)
    return np.concatenate([o1, o2, o3]), np.concatenate([exp[:54], exp[54:108], exp[162:216]]), 1e-6


def case_mlp_fp32_1layer_512(be):
    # test/Integration/mlp-fp32-1layer-512.mlir:8-20: relu(x[128x256] . W[256x512] + bias), all ones => 257; first row
    # printed. As the fused op of the layer (beta_0, add bcast_col_in0, relu)
    x, W, b = const(F32, (128, 256)), const(F32, (256, 512)), const(F32, (512,))
    y = np.zeros(128 * 512, np.float32)
    be.fused_brgemm(F32, 128, 512, 256, 256, 512, 512, 0, 0, 4, 0, 5, 4, 1, x, 0, W, 0, y, 0, b, 0, 1)
    row = golden()["mlp_fp32_1layer_512"]["raw_checks"][0]["values"]
    assert len(row) == 512
    return y[:512], np.array(row, np.float32), 0.0


def case_tpp_brgemm_f32(be):
    # test/Integration/tpp-brgemm.mlir:12-16: linalg.batch_reduce_matmul with ONE batch element, 1x4x8 . 1x8x4 into a zero
    # 4x4; the IR check expects xsmm_gemm_invoke -> gemm [4,4,8,8,4,4]
    g = golden()["tpp_brgemm_f32"]
    A, B, C = np.array(g["A"], np.float32), np.array(g["B"], np.float32), np.zeros(16, np.float32)
    be.gemm(F32, 4, 4, 8, 8, 4, 4, 0, A, 0, B, 0, C, 0)
    return C, np.array(g["expected"], np.float32), 6e-3   # 5 printed digits of values around 5e2


def case_simple_gemm_f32(be):
    # test/Integration/simple-gemm.mlir:5-11: kernel arguments filled by tpp-run's default init (all 1.0, C included):
    # gemm [4,4,8,8,4,4] accumulating => 9
    A, B, C = const(F32, (4, 8)), const(F32, (8, 4)), const(F32, (16,))
    be.gemm(F32, 4, 4, 8, 8, 4, 4, 0, A, 0, B, 0, C, 0)
    return C, np.array(golden()["simple_gemm_f32"]["expected"], np.float32), 0.0


def case_packed_matmul_f32(be):
    # test/Integration/packed-matmul.mlir:25-49: C = bias broadcast over rows; C += A(4x8) . B(8x16); relu(C) in place.
    # (1) unpacked: unary identity [4,16,16,16] (bcast_col) -> gemm [4,16,8,8,16,16] -> unary relu [4,16,16,16];
    # (2) the third RUN line, -pack-matmul="block-factors=2,2,2": operands block-packed to [M/2][K/2][2][2],
    #     [N/2][K/2][2][2] (relayout done here in numpy), per 2x2 output block the same three ops with a BRGEMM over the
    #     4 K blocks: brgemm [2,2,2,2,2,2,4,4]. Both must print the same tensor.
    g = golden()["packed_matmul_f32"]
    A, B, bias = (np.array(g[x], np.float32) for x in ("A", "B", "bias"))
    C = np.zeros(64, np.float32)
    be.unary(1, F32, 4, 16, 16, 16, 4, bias, 0, C, 0)
    be.gemm(F32, 4, 16, 8, 8, 16, 16, 0, A, 0, B, 0, C, 0)
    be.unary(5, F32, 4, 16, 16, 16, 0, C, 0, C, 0)
    Ap = np.ascontiguousarray(A.reshape(2, 2, 4, 2).transpose(0, 2, 1, 3)).reshape(-1)    # [i][k][ii][kk]
    Bp = np.ascontiguousarray(B.reshape(4, 2, 8, 2).transpose(2, 0, 1, 3)).reshape(-1)    # [j][k][kk][jj]
    Cp = np.zeros(64, np.float32)                                                         # [i][j][ii][jj]
    for i in range(2):
        for j in range(8):
            off = (i * 8 + j) * 4
            be.unary(1, F32, 2, 2, 2, 2, 4, bias, 2 * j, Cp, off)
            be.brgemm(F32, 2, 2, 2, 2, 2, 2, 4, 4, 0, Ap, i * 16, Bp, j * 16, Cp, off, 4)
            be.unary(5, F32, 2, 2, 2, 2, 0, Cp, off, Cp, off)
    unpacked = Cp.reshape(2, 8, 2, 2).transpose(0, 2, 1, 3).reshape(-1)
    exp = np.array(g["expected"], np.float32)
    return np.concatenate([C, unpacked]), np.concatenate([exp, exp]), 6e-3   # 5 printed digits of values up to 7e2


def case_xsmm_path_add_f32(be):
    # test/Integration/tpp-run-xsmm-path.mlir:8-38: binary add [2,2,2,2,2] of a 1.0 fill and a 2.0 fill, written over the
    # second operand (outs(%arg1)) => 3
    a, b = const(F32, (4,), 1.0), const(F32, (4,), 2.0)
    be.binary(1, F32, 2, 2, 2, 2, 2, 0, a, 0, b, 0, b, 0)
    return b, np.array(golden()["xsmm_path_add_f32"]["expected"], np.float32), 0.0


def case_matmul_tpp_with_print_f32(be):
    # test/Integration/matmul-tpp-with-print.mlir:33-59: fills 1.0 / 2.0 / 0.0, then gemm [4,4,8,8,4,4]: the zero fill
    # of C as its own xsmm zero op [4,4,4,4] followed by the accumulating gemm (the pair fuseZeroWithGemmOrBrgemm folds,
    # ConvertLinalgToXsmm.cpp:962-993) => 16
    A, B, C = const(F32, (4, 8), 1.0), const(F32, (8, 4), 2.0), const(F32, (16,), 7.0)
    be.unary(2, F32, 4, 4, 4, 4, 0, C, 0, C, 0)
    be.gemm(F32, 4, 4, 8, 8, 4, 4, 0, A, 0, B, 0, C, 0)
    return C, np.array(golden()["matmul_tpp_with_print_f32"]["expected"], np.float32), 0.0


def case_result_out_arg_f32(be):
    # test/Integration/result-out-arg.mlir:9-48: out (a kernel argument, default-initialised to 1.0) is zeroed, then
    # gemm [2,2,2,2,2,2] of two 2x2 literals => ( 4, 5 ), ( 10, 11 )
    g = golden()["result_out_arg_f32"]
    A, B, C = np.array(g["A"], np.float32), np.array(g["B"], np.float32), const(F32, (4,))
    be.unary(2, F32, 2, 2, 2, 2, 0, C, 0, C, 0)
    be.gemm(F32, 2, 2, 2, 2, 2, 2, 0, A, 0, B, 0, C, 0)
    return C, np.array(g["expected"], np.float32), 0.0


def case_tpp_pack_unpack_f32(be):
    # test/Integration/tpp-pack-unpack.mlir:3-91, tile by tile as identity copies (what tensor.pack / unpack lower to,
    # LowerPacksAndUnpacks.cpp:143-250): pack1 4x4 -> [2][2][2][2] (tiles of 2x2: identity [2,2,4,2]); pack2 1x2x2x4 ->
    # [1][2][2][2][2] (outer_dims_perm [0,3,1,2], inner tile 2 of the last dim: per half of the last dim a 4x2 copy,
    # identity [4,2,4,2]); unpack1 / unpack2 are the inverses. (pack3 is a CHECK-NOT in the test.)
    g = golden()["tpp_pack_unpack_f32"]
    outs = []
    src, dst = np.array(g["pack1_in"], np.float32), np.zeros(16, np.float32)
    for i in range(2):
        for j in range(2):
            be.unary(1, F32, 2, 2, 4, 2, 0, src, i * 8 + j * 2, dst, (i * 2 + j) * 4)
    outs.append(dst)
    src, dst = np.array(g["pack2_in"], np.float32), np.zeros(16, np.float32)
    for dd in range(2):
        be.unary(1, F32, 4, 2, 4, 2, 0, src, dd * 2, dst, dd * 8)
    outs.append(dst)
    src, dst = np.array(g["unpack1_in"], np.float32), np.zeros(16, np.float32)
    for i in range(2):
        for j in range(2):
            be.unary(1, F32, 2, 2, 2, 4, 0, src, (i * 2 + j) * 4, dst, i * 8 + j * 2)
    outs.append(dst)
    src, dst = np.array(g["unpack2_in"], np.float32), np.zeros(16, np.float32)
    for dd in range(2):
        be.unary(1, F32, 4, 2, 2, 4, 0, src, dd * 8, dst, dd * 2)
    outs.append(dst)
    exp = np.array(g["expected_all"], np.float32)
    assert exp.size == 64
    return np.concatenate(outs), exp, 0.0


def case_tiling_add_f32(be):
    # test/Integration/tiling-add.mlir:14-24: B += A on 32x16 dense literals -> binary add [32,16,16,16,16], result over
    # the second operand
    g = golden()["tiling_add_f32"]
    A, B = np.array(g["A"], np.float32), np.array(g["B"], np.float32)
    be.binary(1, F32, 32, 16, 16, 16, 16, 0, A, 0, B, 0, B, 0)
    return B, np.array(g["expected"], np.float32), 2e-5   # up to 4 printed digits, f32 sums of decimal literals


def case_tiling_relu_f32(be):
    # test/Integration/tiling-relu.mlir:14-24: relu in place on a 32x16 dense literal -> unary relu [32,16,16,16]
    g = golden()["tiling_relu_f32"]
    x = np.array(g["input"], np.float32)
    be.unary(5, F32, 32, 16, 16, 16, 0, x, 0, x, 0)
    return x, np.array(g["expected"], np.float32), 1e-6


def case_mlp_single_layer_bf16(be):
    # test/BF16/Integration/mlp-single-layer-bf16.mlir:11-55: C(128x512) = bias broadcast (identity bcast_col); C += x . W
    # with x = 128x128x2 (collapsed: 128x256) and VNNI-2 weights 128x512x2 -> gemm [128,512,256,256,512,512] (vnni_b);
    # relu in place; all ones: 1 + 256 = 257 -> 256 in bf16 (expected 256, threshold 1.0)
    g = golden()["mlp_single_layer_bf16"]
    x, W, bias = const(BF16, (128, 256)), const(BF16, (128, 512, 2)), const(BF16, (512,))
    C = np.zeros(128 * 512, np.uint16)
    be.unary(1, BF16, 128, 512, 512, 512, 4, bias, 0, C, 0)
    be.gemm(BF16, 128, 512, 256, 256, 512, 512, 2048, x, 0, W, 0, C, 0)
    be.unary(5, BF16, 128, 512, 512, 512, 0, C, 0, C, 0)
    return to_f32(BF16, C), np.full(C.size, g["expected_fill"], np.float32), g["threshold"]


def case_broadcast_row_1d_f32_seed123(be):
    # test/Integration/broadcast-row-1d.mlir:5-21 (--seed 123): dispatch (1,1,4,2,1,2,2) = identity, f32, [4,2,1,2],
    # bcast_row; kernel args filled in order (arg0 4x1, arg1 4x2) from one normal generator, arg1 overwritten
    gen = oracle.TensorInit("normal", F32, 123)
    src, dst = gen.fill(4, 1), gen.fill(4, 2)
    be.unary(1, F32, 4, 2, 1, 2, 2, src, 0, dst, 0)
    return dst.reshape(-1), np.array(golden()["broadcast_row_1d_f32_seed123"]["expected"], np.float32), 1e-6   # 6 printed digits


def _to_blocks(be, src, rows, cols, dst):
    # "to-block-layout" of relayout-gemm.mlir:9-18 / relayout-more-interesting.mlir:8-17 with 2x2 blocks: block (i,j) of
    # a rows x cols matrix -> dst[i][j][2][2], one identity copy [2,2,cols,2] per block
    for i in range(rows // 2):
        for j in range(cols // 2):
            be.unary(1, F32, 2, 2, cols, 2, 0, src, i * 2 * cols + j * 2, dst, (i * (cols // 2) + j) * 4)


def case_relayout_gemm_f32(be):
    # test/Integration/relayout-gemm.mlir:30-95: A(6x8), B(8x16), C(6x16, zero) relaid out to 2x2 blocks, per output
    # block (p1,p2) and K block r1 a gemm [2,2,2,2,2,2] accumulating into the block, then "from-block-layout"
    # (identity [2,2,2,16] per block) back into C
    g = golden()["relayout_gemm_f32"]
    A, B, C = np.array(g["A"], np.float32), np.array(g["B"], np.float32), np.zeros(96, np.float32)
    Ab, Bb, Cb = np.zeros(48, np.float32), np.zeros(128, np.float32), np.full(96, 7.0, np.float32)
    _to_blocks(be, C, 6, 16, Cb)
    _to_blocks(be, A, 6, 8, Ab)
    _to_blocks(be, B, 8, 16, Bb)
    for p1 in range(3):
        for p2 in range(8):
            for r1 in range(4):
                be.gemm(F32, 2, 2, 2, 2, 2, 2, 0, Ab, (p1 * 4 + r1) * 4, Bb, (r1 * 8 + p2) * 4, Cb, (p1 * 8 + p2) * 4)
    for i in range(3):
        for j in range(8):
            be.unary(1, F32, 2, 2, 2, 16, 0, Cb, (i * 8 + j) * 4, C, i * 32 + j * 2)
    return C, np.array(g["expected"], np.float32), 6e-3   # 5 printed digits of values up to 7e2


def case_relayout_block_copy_f32(be):
    # test/Integration/relayout-more-interesting.mlir:22-105: (1) 6x16 -> [3][8][2][2] by block copies, (2) linalg.copy of
    # the whole matrix (identity [6,16,16,16]) viewed as [3][8][2][2] by a pure reshape - the two printed tensors differ
    g = golden()["relayout_block_copy_f32"]
    d = np.array(g["input"], np.float32)
    blocked, copied = np.zeros(96, np.float32), np.zeros(96, np.float32)
    _to_blocks(be, d, 6, 16, blocked)
    be.unary(1, F32, 6, 16, 16, 16, 0, d, 0, copied, 0)
    exp = np.array(g["expected_all"], np.float32)
    assert exp.size == 192
    return np.concatenate([blocked, copied]), exp, 0.0


def case_smoke_matmul_f32(be):
    # test/Integration/smoke.mlir:13-59: C (dense<0.0>) += A(4x8) . B(8x4): gemm [4,4,8,8,4,4]
    g = golden()["smoke_matmul_f32"]
    A, B, C = np.array(g["A"], np.float32), np.array(g["B"], np.float32), np.zeros(16, np.float32)
    be.gemm(F32, 4, 4, 8, 8, 4, 4, 0, A, 0, B, 0, C, 0)
    return C, np.array(g["expected"], np.float32), 6e-3   # 5 printed digits of values up to 5e2


def _conv_as_gemms(be, H, W, KH, KW, stride):
    # test/Integration/conv-to-matmul.mlir:28-45 after -rewrite-conv-to-matmul-or-brgemm (IR: linalg.matmul): NHWC image
    # 1xHxWx3 (value = channel index), HWCF filter KHxKWx3x8 (value = filter index), output 1xOHxOWx8 preloaded with the
    # filter index; per (oh, kh, kw) one accumulating gemm over an output row: [OW, 8, 3, lda = stride*3, 8, 8] on the
    # image window that starts at pixel (oh*stride + kh, kw)
    OH, OW = (H - KH) // stride + 1, (W - KW) // stride + 1
    img = np.tile(np.arange(3, dtype=np.float32), H * W)
    flt = np.tile(np.arange(8, dtype=np.float32), KH * KW * 3)
    out = np.tile(np.arange(8, dtype=np.float32), OH * OW)
    for oh in range(OH):
        for kh in range(KH):
            for kw in range(KW):
                be.gemm(F32, OW, 8, 3, stride * 3, 8, 8, 0, img, ((oh * stride + kh) * W + kw) * 3,
                        flt, (kh * KW + kw) * 24, out, oh * OW * 8)
    return out


def case_conv_to_matmul_f32(be):
    # test/Integration/conv-to-matmul.mlir:47-152: 1x1 filter on 4x4 (=> 4 f), 3x3 filter on 5x5 (=> 28 f), the same with
    # stride 2 (=> 28 f on a 2x2 output)
    outs = [_conv_as_gemms(be, 4, 4, 1, 1, 1), _conv_as_gemms(be, 5, 5, 3, 3, 1), _conv_as_gemms(be, 5, 5, 3, 3, 2)]
    exp = np.array(golden()["conv_to_matmul_f32"]["expected_all"], np.float32)
    assert exp.size == 128 + 72 + 32
    return np.concatenate(outs), exp, 0.0


def case_broadcast_col_2d_f32_seed123(be):
    # test/Integration/broadcast-2d.mlir:7-15 (--seed 123): an 8-vector broadcast over the rows of a 2x4x8 tensor ->
    # identity [8, 8, 8, 8] with bcast_col; kernel args filled in order (arg0 8, arg1 2x4x8) from one normal generator
    gen = oracle.TensorInit("normal", F32, 123)
    src, dst = gen.fill(8), gen.fill(2, 4, 8)
    be.unary(1, F32, 8, 8, 8, 8, 4, src, 0, dst, 0)
    return dst.reshape(-1), np.array(golden()["broadcast_col_2d_f32_seed123"]["expected"], np.float32), 1e-6


def case_tpp_add_slices_f32(be):
    # test/Integration/tpp-add.mlir:23-71: b0 = rows of 0..31 (56x32); per (i, j) of a 2x56x56x32 tensor the slice
    # [i, j, :, :] = b0 + b0: binary add [56,32,32,32,32] with the output at offset (i*56 + j) * 56*32; every one of the
    # 6272 printed rows is 0, 2, ..., 62
    g = golden()["tpp_add_slices_f32"]
    b0 = np.ascontiguousarray(np.tile(np.arange(32, dtype=np.float32), (56, 1)))
    out = np.full(2 * 56 * 56 * 32, -1.0, np.float32)
    for i in range(2):
        for j in range(56):
            be.binary(1, F32, 56, 32, 32, 32, 32, 0, b0, 0, b0, 0, out, (i * 56 + j) * 56 * 32)
    assert g["repeat"] * len(g["row"]) == out.size
    return out, np.tile(np.array(g["row"], np.float32), g["repeat"]), 0.0


def case_fill_subviews_f32(be):
    # test/Integration/subview-on-tensor.mlir:11-57: linalg.fill of the 3x3 subviews [0,0] and [1,1] of a zero 2x2x3x3
    # tensor with 5.0 / 6.0 -> identity [3,3,1,3] with bcast_scalar, the scalar passed by value
    # (xsmm_unary_scalar_invoke), outputs at offsets 0 and 27
    g = golden()["fill_subviews_f32"]
    A = np.zeros(36, np.float32)
    for value, slot in zip(g["fills"], (0, 1)):
        be.unary_scalar(1, F32, 3, 3, 1, 3, 8, value, A, (slot * 2 + slot) * 9)
    return A, np.array(g["expected"], np.float32), 0.0


CASES = {name[len("case_"):]: fn for name, fn in sorted(globals().items()) if name.startswith("case_")}
