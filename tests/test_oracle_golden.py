"""Pins the CPU oracle against every known-answer vector the reference's own tests hold
for this path. CPU only."""
import numpy as np
import pytest

import oracle
from golden_cases import CASES


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_reference_vector(name, oracle_backend):
    got, want, tol = CASES[name](oracle_backend)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=0, atol=tol, err_msg=name)


def test_bf16_rounding_is_rne():
    # ties to even, and the 257 -> 256 case the reference relies on (xsmm-ternary-bf16.mlir)
    x = np.array([257.0, 258.0, 259.0, 1.0 + 2.0 ** -8, 1.0 + 3 * 2.0 ** -8, -257.0], np.float32)
    got = oracle.bf16_to_f32(oracle.f32_to_bf16(x))
    np.testing.assert_array_equal(got, np.array([256.0, 258.0, 260.0, 1.0, 1.0 + 2.0 ** -6, -256.0], np.float32))


def test_tensor_init_streams_are_sequential():
    # one generator per (type, dtype, seed), consumed tensor after tensor (TensorInit.cpp:75-86)
    g1 = oracle.TensorInit("normal", oracle.F32, 123)
    a, b = g1.fill(8), g1.fill(8)
    g2 = oracle.TensorInit("normal", oracle.F32, 123)
    ab = g2.fill(16)
    np.testing.assert_array_equal(np.concatenate([a, b]), ab)
    assert (ab >= 0).all() and (ab <= 1).all()
    c = oracle.TensorInit("cont", oracle.F32).fill(4)
    np.testing.assert_array_equal(c, np.array([0, 0.25, 0.5, 0.75], np.float32))
    s = oracle.TensorInit("simple", oracle.F32).fill(4)
    np.testing.assert_allclose(s, [0.3, 0.6, 0.9, 0.3])


def test_oracle_thread_count_does_not_change_results():
    rng = np.random.default_rng(0)
    A = oracle.f32_to_bf16(rng.random((4, 40, 24), dtype=np.float32))
    B = oracle.f32_to_bf16(rng.random((4, 24, 56), dtype=np.float32))
    outs = []
    for t in (1, 4):
        oracle.set_num_threads(t)
        C = np.zeros((40, 56), np.uint16)
        oracle.brgemm(2, 40, 56, 24, 24, 56, 56, 40 * 24, 24 * 56, 4, A, B, C, 4)
        outs.append(C)
    oracle.set_num_threads(0)
    np.testing.assert_array_equal(outs[0], outs[1])


def test_oracle_vnni_equals_flat():
    rng = np.random.default_rng(1)
    m, n, k, batch = 9, 10, 12, 3
    A = oracle.f32_to_bf16(rng.random((batch, m, k), dtype=np.float32))
    Bf = rng.random((batch, k, n), dtype=np.float32)
    B = oracle.f32_to_bf16(Bf)
    Bv = np.ascontiguousarray(B.reshape(batch, k // 2, 2, n).transpose(0, 1, 3, 2))
    C0, C1 = np.zeros((m, n), np.uint16), np.zeros((m, n), np.uint16)
    oracle.brgemm(2, m, n, k, k, n, n, m * k, k * n, 4, A, B, C0, batch)
    oracle.brgemm(2, m, n, k, k, n, n, m * k, k * n, 4 | 2048, A, Bv, C1, batch)
    np.testing.assert_array_equal(C0, C1)


@pytest.mark.parametrize("native", [False, True])
def test_fast_cpu_kernel_matches_oracle(native):
    """The vectorised kernel that bench.py times as the CPU arm computes the same operator as the oracle
    (native=True: the -march=native build, i.e. the vdpbf16ps microkernel where the CPU has AVX512-BF16)."""
    oracle.use_native(native)
    try:
        _check_fast_kernel()
    finally:
        oracle.use_native(False)
        oracle.lib()


def _check_fast_kernel():
    rng = np.random.default_rng(5)
    m, n, k, batch = 64, 96, 128, 3
    A = oracle.f32_to_bf16(rng.uniform(-1, 1, (batch, m, k)).astype(np.float32))
    B = oracle.f32_to_bf16(rng.uniform(-1, 1, (batch, k, n)).astype(np.float32))
    bias = oracle.f32_to_bf16(rng.uniform(-1, 1, (n,)).astype(np.float32))
    for gflags in (4, 0):
        C0 = oracle.f32_to_bf16(rng.uniform(-1, 1, (m, n)).astype(np.float32))
        c_ref, c_fast = C0.copy(), C0.copy()
        oracle.fused_brgemm(2, m, n, k, k, n, n, m * k, k * n, gflags, 0, 5, 4, 1, A, B, c_ref, bias, batch)
        assert oracle.fused_brgemm_fast(2, m, n, k, k, n, n, m * k, k * n, gflags, 5, 4, 1, A, B, c_fast, bias, batch)
        r, f = oracle.bf16_to_f32(c_ref), oracle.bf16_to_f32(c_fast)
        np.testing.assert_allclose(f, r, rtol=1e-2, atol=1e-2 * np.abs(r).max())
    # unsupported shapes are refused, not mis-computed
    assert not oracle.fused_brgemm_fast(2, 7, n, k, k, n, n, 0, 0, 4, 5, 4, 1, A, B, c_fast, bias, 1)
    # the AMX-BF16 tile kernel (VNNI-2 packed B), where the build and the host have AMX
    if oracle.has_amx():
        Bv = np.zeros((batch, k // 2, n, 2), np.uint16)
        for b in range(batch):
            oracle.unary(28, 2, k, n, n, n, 0, B[b], Bv[b])
        for gflags in (4, 0):
            C0 = oracle.f32_to_bf16(rng.uniform(-1, 1, (m, n)).astype(np.float32))
            c_ref, c_amx = C0.copy(), C0.copy()
            oracle.fused_brgemm(2, m, n, k, k, n, n, m * k, k * n, gflags, 0, 5, 4, 1, A, B, c_ref, bias, batch)
            assert oracle.fused_brgemm_amx(2, m, n, k, k, n, n, m * k, k * n, gflags | 2048, 5, 4, 1, A, Bv, c_amx, bias, batch)
            r, f = oracle.bf16_to_f32(c_ref), oracle.bf16_to_f32(c_amx)
            np.testing.assert_allclose(f, r, rtol=1e-2, atol=1e-2 * np.abs(r).max())
        assert not oracle.fused_brgemm_amx(2, m, n, k, k, n, n, 0, 0, 4, 5, 4, 1, A, B, c_fast, bias, 1)   # flat B: refused


def test_vnni4_layout_of_the_oracle():
    """VNNI-4 has no golden vector in the reference's tests (parity unpinned for this layout): the oracle is checked
    against the layout definition itself, B[K/4][N][4] (lib/TPP/Transforms/Utils/VNNIUtils.cpp:75-78 with factor 4), and
    the VNNI-4 BRGEMM against the flat one on the same numbers."""
    rng = np.random.default_rng(44)
    k, n, m = 16, 24, 8
    B = rng.integers(0, 1 << 15, size=(k, n), dtype=np.uint16)
    P = np.zeros((k // 4, n, 4), np.uint16)
    assert oracle.unary(32, 2, k, n, n, n, 0, B, P) is None or True
    np.testing.assert_array_equal(P, B.reshape(k // 4, 4, n).transpose(0, 2, 1))
    back = np.zeros((k, n), np.uint16)
    oracle.unary(1032, 2, k, n, n, n, 0, P, back)
    np.testing.assert_array_equal(back, B)
    A = oracle.f32_to_bf16(rng.uniform(-1, 1, (m, k)).astype(np.float32))
    Bf = oracle.f32_to_bf16(rng.uniform(-1, 1, (k, n)).astype(np.float32))
    Bv = np.ascontiguousarray(Bf.reshape(k // 4, 4, n).transpose(0, 2, 1))
    c_flat, c_v4 = np.zeros((m, n), np.uint16), np.zeros((m, n), np.uint16)
    oracle.brgemm(2, m, n, k, k, n, n, 0, 0, 4, A, Bf, c_flat, 1)
    oracle.set_vnni_factor(4)
    try:
        oracle.brgemm(2, m, n, k, k, n, n, 0, 0, 4 | 2048, A, Bv, c_v4, 1)
    finally:
        oracle.set_vnni_factor(2)
    np.testing.assert_array_equal(c_flat, c_v4)


def test_xsmm_semantics_agree_with_the_loops_path_like_the_reference_checks():
    """test/BF16/Integration/vnni-xsmm-vs-loops.mlir:1-13: the reference runs
    `mlir-gen --kernel=const --bias --relu --seed=123 --batch=16 --layers=16,16 --tiles=16,16,16 --float-type=bf16`
    once through the xsmm path and once with -linalg-to-loops and accepts `fpcmp -r 0.01` between the two printed
    tensors. The loops path is independent of libxsmm: the linalg.generic bodies are plain arith.mulf / arith.addf in
    bf16 (every product and every partial sum rounded to bf16), then bias add and max(x, 0) in bf16. Restated here in
    numpy and compared with the oracle (f32 accumulation, one rounding) under fpcmp's rule |a / b - 1| <= 0.01
    (tools/fpcmp/fpcmp.c:198-205) - the same independent cross-check of the oracle's semantics the reference applies to
    libxsmm. Data: weights from TensorInit(normal, seed 123), bias from the seed mlir-gen draws next (rand() after
    srand(123): tools/mlir-gen/MLIRGen.cpp:131-134, 256-259, 812-818), input from tpp-run's own seed-123 generator."""
    import ctypes

    libc = ctypes.CDLL("libc.so.6")
    libc.srand(123)
    bias_seed = libc.rand()
    m = n = k = 16
    W = oracle.TensorInit("normal", oracle.BF16, 123).fill(k, n)
    bias = oracle.TensorInit("normal", oracle.BF16, bias_seed).fill(n)
    x = oracle.TensorInit("normal", oracle.BF16, 123).fill(m, k)

    got = np.zeros((m, n), np.uint16)
    oracle.fused_brgemm(2, m, n, k, k, n, n, 0, 0, 4, 0, 5, 4, 1, x, W, got, bias, 1)
    xsmm_path = oracle.bf16_to_f32(got)

    def r(a):   # one bf16 rounding
        return oracle.bf16_to_f32(oracle.f32_to_bf16(np.ascontiguousarray(a, dtype=np.float32)))

    xf, wf, bf = oracle.bf16_to_f32(x), oracle.bf16_to_f32(W), oracle.bf16_to_f32(bias)
    acc = np.zeros((m, n), np.float32)
    for c in range(k):   # reduction loop of the generic: out = out + in0 * in1, both ops in bf16
        acc = r(acc + r(np.outer(xf[:, c], wf[c, :])))
    loops_path = np.maximum(r(acc + bf[None, :]), 0.0)

    assert (loops_path > 0).any()
    both_zero = (xsmm_path == 0) & (loops_path == 0)
    ratio = np.where(both_zero, 1.0, xsmm_path / np.where(loops_path == 0, 1e-30, loops_path))
    assert np.abs(ratio - 1.0).max() <= 0.01, np.abs(ratio - 1.0).max()
