"""GPU parity tests: the CUDA path, called through the C-ABI, against the CPU oracle and the
reference's known-answer vectors. Tolerances (BASELINE.json north_star): bf16 1e-2 rel,
f32 1e-5 rel, bit-exact for data movement (identity / zero / transpose / VNNI)."""
import os

import numpy as np
import pytest

import oracle
from backends import BF16, F32, AbiBackend, OracleBackend, np_dtype
from golden_cases import CASES

pytestmark = pytest.mark.gpu

BF16_RTOL = 1e-2
F32_RTOL = 1e-5


def rnd(rng, dtype, shape, lo=-1.0, hi=1.0):
    a = rng.uniform(lo, hi, size=shape).astype(np.float32)
    return a if dtype == F32 else oracle.f32_to_bf16(a)


def as_f32(dtype, a):
    return a.astype(np.float32) if dtype == F32 else oracle.bf16_to_f32(a)


def assert_close(dtype, got, want, scale=None):
    g, w = as_f32(dtype, got).astype(np.float64), as_f32(dtype, want).astype(np.float64)
    rtol = F32_RTOL if dtype == F32 else BF16_RTOL
    s = np.abs(w).max() if scale is None else scale
    np.testing.assert_allclose(g, w, rtol=rtol, atol=rtol * max(s, 1e-30) * 0.5)


@pytest.fixture(scope="module")
def dev():
    return AbiBackend("device")


@pytest.fixture(scope="module")
def orc():
    return OracleBackend()


# ---- 1. the reference's own known-answer tests, through the ABI ---------------------------
@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_vectors_device(name, dev):
    got, want, tol = CASES[name](dev)
    np.testing.assert_allclose(got, want, rtol=0, atol=tol, err_msg=name)


@pytest.mark.parametrize("placement", ["host", "mirror"])
@pytest.mark.parametrize("name", ["brgemm_f32_ones", "brgemm_bf16_vnni", "fused_bf16_vnni", "fused_f32_seed123",
                                  "transpose_f32", "vnni2_pack_chain", "unary_relu_bf16", "binary_div_f32",
                                  "strided_gemm1", "strided_brgemm", "mlp_all_ones_bf16"])
def test_golden_vectors_host_pointers(name, placement):
    got, want, tol = CASES[name](AbiBackend(placement))
    np.testing.assert_allclose(got, want, rtol=0, atol=tol, err_msg=f"{name}/{placement}")


# ---- 2. elementwise TPPs vs oracle --------------------------------------------------------
UNARY_SHAPES = [(1, 1), (3, 5), (32, 32), (37, 129), (256, 1024), (1, 4096), (513, 8)]


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("kind", [1, 2, 5])
@pytest.mark.parametrize("flags", [0, 2, 4, 8])
@pytest.mark.parametrize("shape", UNARY_SHAPES)
def test_unary_eltwise_bit_exact(dtype, kind, flags, shape, dev, orc):
    m, n = shape
    rng = np.random.default_rng(m * 1000 + n + kind)
    pad = 0 if (m * n) % 2 else 8  # exercise ld > n and the vector / scalar paths
    ldo = n + pad
    if flags == 0:
        ldi, inp = n + pad, rnd(rng, dtype, (m, n + pad))
    elif flags == 2:
        ldi, inp = 1, rnd(rng, dtype, (m,))
    elif flags == 4:
        ldi, inp = n, rnd(rng, dtype, (n,))
    else:
        ldi, inp = 1, rnd(rng, dtype, (1,))
    sentinel = rnd(rng, dtype, (m, ldo))
    out_g, out_o = sentinel.copy(), sentinel.copy()
    dev.unary(kind, dtype, m, n, ldi, ldo, flags, inp, 0, out_g, 0)
    orc.unary(kind, dtype, m, n, ldi, ldo, flags, inp, 0, out_o, 0)
    np.testing.assert_array_equal(out_g.view(np.uint32 if dtype == F32 else np.uint16),
                                  out_o.view(np.uint32 if dtype == F32 else np.uint16))


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_unary_in_place_and_offsets(dtype, dev, orc):
    rng = np.random.default_rng(7)
    buf = rnd(rng, dtype, (4, 64, 48))
    g, o = buf.copy(), buf.copy()
    for be, b in ((dev, g), (orc, o)):
        be.unary(5, dtype, 64, 48, 48, 48, 0, b, 2 * 64 * 48, b, 2 * 64 * 48)  # relu(x, x) on the 3rd slice
    np.testing.assert_array_equal(g, o)
    assert (g[0] == buf[0]).all() and (g[3] == buf[3]).all()


@pytest.mark.parametrize("dtype", [F32, BF16])
def test_unary_scalar_invoke(dtype, dev, orc):
    for kind, val in ((2, 7.5), (1, 0.3), (5, -2.0), (5, 1.7)):
        g = rnd(np.random.default_rng(0), dtype, (33, 40))
        o = g.copy()
        dev.unary_scalar(kind, dtype, 33, 40, 1, 40, 8, val, g, 0)
        orc.unary_scalar(kind, dtype, 33, 40, 1, 40, 8, val, o, 0)
        np.testing.assert_array_equal(g, o)


BIN_FLAGS = [0, 1, 2, 4, 8, 16, 32, 4 | 2, 1 | 8, 16 | 8]


@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("kind", [1, 2, 3, 4])
@pytest.mark.parametrize("flags", BIN_FLAGS)
@pytest.mark.parametrize("shape", [(3, 3), (64, 64), (37, 129), (256, 1024)])
def test_binary_vs_oracle(dtype, kind, flags, shape, dev, orc):
    m, n = shape
    rng = np.random.default_rng(kind * 7 + flags)

    def operand(bits_row, bits_col, bits_scalar):
        if flags & bits_row:
            return 1, rnd(rng, dtype, (m,), 0.5, 2.0)
        if flags & bits_col:
            return n, rnd(rng, dtype, (n,), 0.5, 2.0)
        if flags & bits_scalar:
            return 1, rnd(rng, dtype, (1,), 0.5, 2.0)
        return n, rnd(rng, dtype, (m, n), 0.5, 2.0)

    ldl, lhs = operand(1, 4, 16)
    ldr, rhs = operand(2, 8, 32)
    g, o = np.zeros((m, n), np_dtype(dtype)), np.zeros((m, n), np_dtype(dtype))
    dev.binary(kind, dtype, m, n, ldl, ldr, n, flags, lhs, 0, rhs, 0, g, 0)
    orc.binary(kind, dtype, m, n, ldl, ldr, n, flags, lhs, 0, rhs, 0, o, 0)
    if kind == 4:  # division: the GPU's IEEE div and the CPU's agree to the last bit for f32; allow 1 ulp in bf16
        assert_close(dtype, g, o)
    else:
        np.testing.assert_array_equal(g, o)


def test_binary_bias_add_in_place(dev, orc):
    # the unfused MLP epilogue: binary add(bias[bcast_col_in0], C, C) then relu(C, C)
    rng = np.random.default_rng(3)
    bias, C = rnd(rng, BF16, (1024,)), rnd(rng, BF16, (256, 1024))
    g, o = C.copy(), C.copy()
    for be, c in ((dev, g), (orc, o)):
        be.binary(1, BF16, 256, 1024, 1024, 1024, 1024, 4, bias, 0, c, 0, c, 0)
        be.unary(5, BF16, 256, 1024, 1024, 1024, 0, c, 0, c, 0)
    np.testing.assert_array_equal(g, o)


# ---- 3. transforms: bit-exact ----------------------------------------------------------------
@pytest.mark.parametrize("dtype", [F32, BF16])
@pytest.mark.parametrize("shape", [(4, 8), (1, 7), (64, 64), (65, 127), (300, 33), (1024, 512)])
def test_transpose_bit_exact(dtype, shape, dev, orc):
    m, n = shape
    rng = np.random.default_rng(m + n)
    inp = rnd(rng, dtype, (m, n + 3))
    g, o = np.zeros((n, m + 5), np_dtype(dtype)), np.zeros((n, m + 5), np_dtype(dtype))
    dev.unary(29, dtype, m, n, n + 3, m + 5, 0, inp, 0, g, 0)
    orc.unary(29, dtype, m, n, n + 3, m + 5, 0, inp, 0, o, 0)
    np.testing.assert_array_equal(g, o)


@pytest.mark.parametrize("shape", [(2, 1), (16, 16), (32, 32), (64, 1000), (130, 72), (1024, 1024)])
@pytest.mark.parametrize("pad", [0, 4, 3])
def test_vnni2_pack_unpack_bit_exact(shape, pad, dev, orc):
    m, n = shape
    rng = np.random.default_rng(m * n + pad)
    inp = rnd(rng, BF16, (m, n + pad))
    ldo = n + pad
    g, o = np.zeros((m // 2, ldo, 2), np.uint16), np.zeros((m // 2, ldo, 2), np.uint16)
    dev.unary(28, BF16, m, n, n + pad, ldo, 0, inp, 0, g, 0)
    orc.unary(28, BF16, m, n, n + pad, ldo, 0, inp, 0, o, 0)
    np.testing.assert_array_equal(g, o)
    # inverse (extension kind): round trip restores the input bits
    back = np.zeros((m, n + pad), np.uint16)
    dev.unary(1028, BF16, m, n, ldo, n + pad, 0, g, 0, back, 0)
    np.testing.assert_array_equal(back[:, :n], inp[:, :n])


def test_vnni2_full_size_roundtrip(dev):
    # BASELINE config 4: 4096 x 4096 bf16; size-independent property: unpack(pack(x)) == x, and the
    # packed tensor is a permutation (same multiset checksum)
    rng = np.random.default_rng(4096)
    x = rng.integers(0, 1 << 16, size=(4096, 4096), dtype=np.uint16)
    p, back = np.zeros((2048, 4096, 2), np.uint16), np.zeros((4096, 4096), np.uint16)
    dev.unary(28, BF16, 4096, 4096, 4096, 4096, 0, x, 0, p, 0)
    dev.unary(1028, BF16, 4096, 4096, 4096, 4096, 0, p, 0, back, 0)
    np.testing.assert_array_equal(back, x)
    assert int(p.astype(np.uint64).sum()) == int(x.astype(np.uint64).sum())
    np.testing.assert_array_equal(p[5, 17], x[10:12, 17])


@pytest.mark.parametrize("shape", [(4, 1), (16, 16), (32, 32), (64, 1000), (132, 72), (1024, 1024)])
@pytest.mark.parametrize("pad", [0, 8, 3])
def test_vnni4_pack_unpack_bit_exact(shape, pad, dev, orc):
    """VNNI-4 ([K][N] -> [K/4][N][4]; mlir-gen --vnni=4, benchmarks/config/omp/mlir-bf16.json:65-125; the transform kinds
    runtime/Xsmm/XsmmRunnerUtils.cpp:45-54 lists): bit-exact against the oracle and against numpy's view of the layout."""
    m, n = shape
    rng = np.random.default_rng(m * n + pad)
    inp = rnd(rng, BF16, (m, n + pad))
    ldo = n + pad
    g, o = np.zeros((m // 4, ldo, 4), np.uint16), np.zeros((m // 4, ldo, 4), np.uint16)
    dev.unary(32, BF16, m, n, n + pad, ldo, 0, inp, 0, g, 0)
    orc.unary(32, BF16, m, n, n + pad, ldo, 0, inp, 0, o, 0)
    np.testing.assert_array_equal(g, o)
    np.testing.assert_array_equal(g[:, :n, :], inp[:, :n].reshape(m // 4, 4, n).transpose(0, 2, 1))
    back = np.zeros((m, n + pad), np.uint16)
    dev.unary(1032, BF16, m, n, ldo, n + pad, 0, g, 0, back, 0)
    np.testing.assert_array_equal(back[:, :n], inp[:, :n])


def test_vnni4_full_size_roundtrip(dev):
    rng = np.random.default_rng(4097)
    x = rng.integers(0, 1 << 16, size=(4096, 4096), dtype=np.uint16)
    p, back = np.zeros((1024, 4096, 4), np.uint16), np.zeros((4096, 4096), np.uint16)
    dev.unary(32, BF16, 4096, 4096, 4096, 4096, 0, x, 0, p, 0)
    dev.unary(1032, BF16, 4096, 4096, 4096, 4096, 0, p, 0, back, 0)
    np.testing.assert_array_equal(back, x)
    np.testing.assert_array_equal(p[5, 17], x[20:24, 17])


@pytest.mark.parametrize("shape", [(32, 32, 32, 32), (64, 48, 96, 3), (256, 512, 256, 4), (256, 1024, 1024, 1)])
def test_brgemm_bf16_vnni4_b(shape, dev, orc, monkeypatch):
    """B operands packed with VNNI factor 4 ([K/4][N][4]): the factor is what libxsmm_cpuid_dot_pack_factor answers
    (TPP_XSMM_VNNI=4), as in the reference, where the dispatch only carries the vnni_b flag."""
    m, n, k, batch = shape
    monkeypatch.setenv("TPP_XSMM_VNNI", "4")
    oracle.set_vnni_factor(4)
    try:
        rng = np.random.default_rng(m + n + k)
        A = rnd(rng, BF16, (batch, m, k))
        Bflat = rnd(rng, BF16, (batch, k, n))
        Bv = np.ascontiguousarray(Bflat.reshape(batch, k // 4, 4, n).transpose(0, 1, 3, 2))
        g, o, ref = (np.zeros((m, n), np.uint16) for _ in range(3))
        dev.brgemm(BF16, m, n, k, k, n, n, m * k, k * n, 4 | 2048, A, 0, Bv, 0, g, 0, batch)
        orc.brgemm(BF16, m, n, k, k, n, n, m * k, k * n, 4 | 2048, A, 0, Bv, 0, o, 0, batch)
        oracle.set_vnni_factor(2)
        orc.brgemm(BF16, m, n, k, k, n, n, m * k, k * n, 4, A, 0, Bflat, 0, ref, 0, batch)   # the flat form of the same math
        np.testing.assert_array_equal(o, ref)
        assert_close(BF16, g, o)
        if m * n * k * batch >= 1 << 21:
            assert dev.kernels[-1] == "vnni4_unpack+brgemm_tc_bf16", dev.kernels[-1]
    finally:
        oracle.set_vnni_factor(2)


# ---- 4. BRGEMM family ---------------------------------------------------------------------------
def run_brgemm_pair(dev, orc, dtype, m, n, k, batch, *, lda=None, ldb=None, ldc=None, sa=None, sb=None, flags=4,
                    vnni=False, fused=None, seed=0, lo=-1.0, hi=1.0):
    """fused = (unary_kind, binary_flags, binary_kind) or None. Returns (gpu C, oracle C, kernel name)."""
    lda, ldb, ldc = lda or k, ldb or n, ldc or n
    sa = m * lda if sa is None else sa
    sb = k * ldb if sb is None else sb
    rng = np.random.default_rng(seed)
    nb = max(batch, 1)
    A = rnd(rng, dtype, ((nb - 1) * sa + m * lda,), lo, hi)
    Bflat = rnd(rng, dtype, ((nb - 1) * sb + k * ldb,), lo, hi)
    B = Bflat
    if vnni:
        # repack every batch's [k][ldb] block as [k/2][ldb][2]
        B = Bflat.copy()
        for b in range(nb):
            blk = Bflat[b * sb:b * sb + k * ldb].reshape(k // 2, 2, ldb)
            B[b * sb:b * sb + k * ldb] = blk.transpose(0, 2, 1).reshape(-1)
        flags |= 2048
    C0 = rnd(rng, dtype, (m * ldc,), lo, hi)
    D = None
    if fused:
        bf = fused[1]
        D = rnd(rng, dtype, (n if bf == 4 else m if bf == 1 else 1 if bf == 16 else m * ldc,), lo, hi)
    g, o = C0.copy(), C0.copy()
    for be, c in ((dev, g), (orc, o)):
        if fused:
            be.fused_brgemm(dtype, m, n, k, lda, ldb, ldc, sa, sb, flags, 0, fused[0], fused[1], fused[2], A, 0, B, 0,
                            c, 0, D, 0, batch)
        else:
            be.brgemm(dtype, m, n, k, lda, ldb, ldc, sa, sb, flags, A, 0, B, 0, c, 0, batch)
    return g, o, dev.kernels[-1]


TC_SHAPES = [
    # m, n, k, batch
    (128, 64, 64, 1), (128, 128, 128, 2), (256, 256, 64, 4), (32, 32, 32, 32), (64, 48, 96, 3),
    (130, 72, 40, 2), (1, 8, 8, 1), (257, 1000, 136, 2), (256, 1024, 1024, 1), (512, 512, 256, 3),
    # wide tiles + split-K over a cluster (128x256 split 4, 128x128 split 4, 128x256 split 2, ragged edges)
    (1024, 1024, 256, 4), (1024, 512, 128, 4), (2048, 1024, 128, 2), (1000, 1000, 264, 3), (1024, 1024, 64, 1),
]


@pytest.mark.parametrize("shape", TC_SHAPES)
@pytest.mark.parametrize("beta0", [True, False])
def test_brgemm_bf16_tensor_core_path(shape, beta0, dev, orc):
    m, n, k, batch = shape
    g, o, kern = run_brgemm_pair(dev, orc, BF16, m, n, k, batch, flags=4 if beta0 else 0, seed=m + n + k)
    assert kern.startswith("brgemm_tc_bf16"), kern
    assert_close(BF16, g, o)


@pytest.mark.parametrize("shape", [(128, 128, 64, 2), (256, 1024, 64, 16), (100, 200, 72, 3)])
def test_brgemm_bf16_padded_leading_dims(shape, dev, orc):
    m, n, k, batch = shape
    g, o, kern = run_brgemm_pair(dev, orc, BF16, m, n, k, batch, lda=k + 8, ldb=n + 16, ldc=n + 24,
                                 sa=m * (k + 8) + 64, sb=k * (n + 16) + 32, seed=11)
    assert kern.startswith("brgemm_tc_bf16"), kern
    assert_close(BF16, g, o)
    # columns >= n of every C row are not written
    gg, oo = g.reshape(m, n + 24), o.reshape(m, n + 24)
    np.testing.assert_array_equal(gg[:, n:], oo[:, n:])


@pytest.mark.parametrize("fused", [(5, 4, 1), (0, 4, 1), (5, 0, 0), (5, 1, 1), (5, 16, 2), (0, 0, 3)])
@pytest.mark.parametrize("shape", [(256, 1024, 1024, 1), (64, 64, 32, 8), (129, 65, 72, 2), (1024, 1024, 128, 4),
                                   (1024, 520, 192, 2)])
def test_fused_brgemm_bf16(fused, shape, dev, orc):
    m, n, k, batch = shape
    ld = {} if n % 8 == 0 else dict(ldb=n + 8 - n % 8, ldc=n + 8 - n % 8)  # TMA strides are multiples of 16 B
    g, o, kern = run_brgemm_pair(dev, orc, BF16, m, n, k, batch, fused=fused, seed=5, **ld)
    assert kern.startswith("brgemm_tc_bf16"), kern
    assert_close(BF16, g, o)
    if fused[0] == 5:
        ldc = ld.get("ldc", n)
        assert (as_f32(BF16, g).reshape(m, ldc)[:, :n] >= 0).all()


def test_fused_brgemm_accumulates_into_c(dev, orc):
    g, o, _ = run_brgemm_pair(dev, orc, BF16, 128, 128, 64, 4, flags=0, fused=(5, 4, 1), seed=9)
    assert_close(BF16, g, o)


def test_cfg3_k64_batch16_equals_k1024_batch1(dev, orc):
    """the strided view (k=64 x batch 16, lda=1024, stride_a=64, stride_b=65536) and the flat
    call (k=1024 x batch 1) are the same math and must agree; both vs oracle."""
    rng = np.random.default_rng(21)
    A, W, bias = rnd(rng, BF16, (256, 1024), 0, 1), rnd(rng, BF16, (1024, 1024), 0, 0.1), rnd(rng, BF16, (1024,))
    outs = []
    for (k, batch, sa, sb) in ((1024, 1, 256 * 1024, 1024 * 1024), (64, 16, 64, 64 * 1024)):
        c = np.zeros((256, 1024), np.uint16)
        dev.fused_brgemm(BF16, 256, 1024, k, 1024, 1024, 1024, sa, sb, 4, 0, 5, 4, 1, A, 0, W, 0, c, 0, bias, 0, batch)
        assert dev.kernels[-1].startswith("brgemm_tc_bf16")
        outs.append(c)
    ref = np.zeros((256, 1024), np.uint16)
    orc.fused_brgemm(BF16, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4, 0, 5, 4, 1, A, 0, W, 0, ref, 0, bias, 0, 1)
    assert_close(BF16, outs[0], ref)
    assert_close(BF16, outs[1], ref)
    np.testing.assert_array_equal(outs[0], outs[1])  # same accumulation order inside the kernel


def test_brgemm_zero_batches(dev, orc):
    # numBatches = 0: C = beta*C (+ post-ops); nothing is read from A/B
    g, o, _ = run_brgemm_pair(dev, orc, BF16, 64, 64, 64, 0, flags=0, fused=(5, 4, 1), seed=2)
    np.testing.assert_array_equal(g, o)
    g, o, _ = run_brgemm_pair(dev, orc, F32, 5, 7, 3, 0, flags=4, seed=2)
    np.testing.assert_array_equal(g, o)


@pytest.mark.parametrize("shape", [(4, 4, 4, 64), (6, 6, 6, 2), (32, 32, 32, 2), (64, 64, 64, 1), (33, 65, 17, 3),
                                   (128, 256, 64, 2)])
@pytest.mark.parametrize("beta0", [True, False])
def test_brgemm_f32_simt(shape, beta0, dev, orc):
    m, n, k, batch = shape
    oracle.set_acc_mode(1)  # f64 truth: summation order is unspecified in the reference
    try:
        g, o, kern = run_brgemm_pair(dev, orc, F32, m, n, k, batch, flags=4 if beta0 else 0, seed=m * k, lo=0, hi=1)
    finally:
        oracle.set_acc_mode(0)
    assert kern.startswith("brgemm_simt_f32"), kern
    assert_close(F32, g, o)


@pytest.mark.parametrize("shape", [(4, 4, 4, 64), (6, 6, 6, 2), (32, 32, 32, 4), (100, 72, 64, 2), (256, 512, 128, 2),
                                   (256, 1024, 1024, 1), (1024, 1024, 64, 16)])
def test_brgemm_bf16_vnni_b(shape, dev, orc):
    m, n, k, batch = shape
    g, o, kern = run_brgemm_pair(dev, orc, BF16, m, n, k, batch, vnni=True, fused=(5, 4, 1), seed=m)
    assert_close(BF16, g, o)
    # VNNI-2 weights reach the tensor cores once the problem is big enough: rewritten in shared memory by the CTA-pair
    # kernel (short reductions) or through one un-interleave pass in front of the flat kernel
    if m * n * k * batch >= 1 << 21:
        assert kern == "vnni2_unpack+brgemm_tc_bf16" or (kern.startswith("brgemm_tc_bf16_256x") and kern.endswith("_vnni2")), kern
    else:
        assert kern.startswith("brgemm_simt_bf16"), kern


@pytest.mark.parametrize("shape", [(256, 1024, 1024, 1), (1024, 1024, 64, 16), (512, 768, 128, 3), (300, 520, 192, 2)])
@pytest.mark.parametrize("native", ["1", "0"])
def test_brgemm_bf16_vnni_b_native_and_unpack_paths_agree(shape, native):
    """VNNI-2 B on the tensor cores, both ways (TPP_XSMM_VNNI_NATIVE is read once per process, hence the subprocess): the
    CTA-pair kernel's in-kernel rewrite (incl. the flag-synchronised split-K variant) and the un-interleave pass + flat
    kernel give the oracle's answer."""
    import subprocess
    import sys

    m, n, k, batch = shape
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = f"""
import sys; sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests')!r})
import numpy as np, oracle, backends
from test_parity_gpu import run_brgemm_pair, assert_close
dev, orc = backends.AbiBackend('device'), backends.OracleBackend()
g, o, kern = run_brgemm_pair(dev, orc, 2, {m}, {n}, {k}, {batch}, vnni=True, fused=(5, 4, 1), seed={m})
assert_close(2, g, o)
print('KERNEL', kern)
"""
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, TPP_XSMM_VNNI_NATIVE=native), capture_output=True,
                         text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    kern = out.stdout.split("KERNEL")[-1].strip()
    if native == "0":
        assert kern == "vnni2_unpack+brgemm_tc_bf16", kern
    else:   # the in-kernel path needs the CTA-pair tiling; shapes the cost model gives to other tilings keep the pass
        assert kern == "vnni2_unpack+brgemm_tc_bf16" or kern.endswith("_vnni2"), kern


def test_brgemm_unaligned_falls_back_to_generic_kernel(dev, orc):
    # lda = 6 elements = 12 bytes: not expressible as a TMA stride -> generic kernel, same answer
    g, o, kern = run_brgemm_pair(dev, orc, BF16, 6, 6, 6, 2, seed=1)
    assert kern.startswith("brgemm_simt_bf16"), kern
    assert_close(BF16, g, o)


def test_gemm_is_brgemm_batch1(dev, orc):
    rng = np.random.default_rng(8)
    A, B = rnd(rng, BF16, (96, 160)), rnd(rng, BF16, (160, 224))
    g, o = np.zeros((96, 224), np.uint16), np.zeros((96, 224), np.uint16)
    dev.gemm(BF16, 96, 224, 160, 160, 224, 224, 4, A, 0, B, 0, g, 0)
    orc.gemm(BF16, 96, 224, 160, 160, 224, 224, 4, A, 0, B, 0, o, 0)
    assert_close(BF16, g, o)


def test_cfg2_full_size_row_samples(dev, orc):
    """BASELINE config 2 at full size: bf16 M=N=K=1024, batch 16 (34.4 GFLOP). The oracle checks row slabs
    (it would need minutes for the whole matrix); a checksum-of-products property covers the rest."""
    rng = np.random.default_rng(1024)
    A = rnd(rng, BF16, (16, 1024, 1024), 0, 1)
    B = rnd(rng, BF16, (16, 1024, 1024), 0, 0.05)
    C = np.zeros((1024, 1024), np.uint16)
    dev.brgemm(BF16, 1024, 1024, 1024, 1024, 1024, 1024, 1 << 20, 1 << 20, 4 | 64 | 128, A, 0, B, 0, C, 0, 16)
    assert dev.kernels[-1].startswith("brgemm_tc_bf16")
    for r0 in (0, 500, 1016):
        ref = np.zeros((8, 1024), np.uint16)
        orc.brgemm(BF16, 8, 1024, 1024, 1024, 1024, 1024, 1 << 20, 1 << 20, 4, A, r0 * 1024, B, 0, ref, 0, 16)
        assert_close(BF16, C[r0:r0 + 8], ref)
    # column-sum property: sum_i C[i][j] == sum_b sum_p (sum_i A[b][i][p]) * B[b][p][j]  (f64), within bf16 rounding
    Af, Bf = as_f32(BF16, A).astype(np.float64), as_f32(BF16, B).astype(np.float64)
    want = np.einsum("bp,bpj->j", Af.sum(axis=1), Bf)
    got = as_f32(BF16, C).astype(np.float64).sum(axis=0)
    np.testing.assert_allclose(got, want, rtol=2e-3)


def test_cfg3_mlp_end_to_end(dev, orc):
    """BASELINE config 3: 3 x fused_brgemm(256 x 1024 x 1024) + bias + relu, TensorInit 'normal' seed 123 data."""
    gen = oracle.TensorInit("normal", BF16, 123)
    Ws, bs = [], []
    for _ in range(3):  # splat constants are replaced first, in op order: W1, b1, W2, b2, W3, b3
        Ws.append(gen.fill(1024, 1024))
        bs.append(gen.fill(1024))
    x = gen.fill(256, 1024)
    xg, xo = x, x
    for W, b in zip(Ws, bs):
        yg, yo = np.zeros((256, 1024), np.uint16), np.zeros((256, 1024), np.uint16)
        dev.fused_brgemm(BF16, 256, 1024, 1024, 1024, 1024, 1024, 256 * 1024, 1 << 20, 4 | 64 | 128, 0, 5, 4, 1, xg, 0,
                         W, 0, yg, 0, b, 0, 1)
        orc.fused_brgemm(BF16, 256, 1024, 1024, 1024, 1024, 1024, 256 * 1024, 1 << 20, 4, 0, 5, 4, 1, xo, 0, W, 0, yo,
                         0, b, 0, 1)
        assert_close(BF16, yg, yo)
        xg, xo = yg, yo


# ---- 5. harness replay + residency modes --------------------------------------------------------
@pytest.mark.parametrize("tiles", [(256, 1024, 1024), (32, 32, 32), (64, 256, 256)])
@pytest.mark.parametrize("vnni", [False, True])
def test_mlp_replay_blocked_layouts(tiles, vnni):
    import torch

    from tpp_mlir_b200 import harness, xsmm

    bn, bk, bc = tiles
    cfg = harness.MlpConfig(batch=256, layers=(1024, 1024, 1024), tiles=tiles, vnni=vnni)
    gen = oracle.TensorInit("normal", BF16, 123)
    Ws = [gen.fill(1024, 1024) for _ in range(2)]
    bs = [gen.fill(1024) for _ in range(2)]
    x = gen.fill(256, 1024)

    def t(a):
        return torch.from_numpy(a.view(np.int16))

    wp = [harness.pack_weight(t(W), bk, bc) for W in Ws]
    if vnni:
        wp = [harness.vnni_pack_weight(w) for w in wp]
    acts = [harness.pack_activation(t(x), bn, bc).cuda()] + [torch.zeros(256 * 1024, dtype=torch.int16).cuda()
                                                               for _ in range(2)]
    r = harness.MlpReplay(cfg, [w.cuda() for w in wp], [t(b).cuda() for b in bs], acts)
    out = r.forward()
    xsmm.sync()
    got = harness.unpack_activation(out.reshape(256 // bn, 1024 // bk, bn, bk)).cpu().numpy().view(np.uint16)
    ref = x
    for W, b in zip(Ws, bs):
        y = np.zeros((256, 1024), np.uint16)
        oracle.fused_brgemm(BF16, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4, 0, 5, 4, 1, ref, W, y, b, 1)
        ref = y
    assert_close(BF16, got, ref)


def _blocked_mlp(tiles, vnni, batch=256, layers=(1024, 1024, 1024), seed=123, n_sets=1, vnni_factor=2):
    """n_sets operand sets of a block-packed MLP (mlir-gen layouts) on the GPU plus the oracle's
    answer for each; returns (cfg, replay objects, expected outputs)."""
    import torch

    from tpp_mlir_b200 import harness

    bn, bk, bc = tiles
    cfg = harness.MlpConfig(batch=batch, layers=layers, tiles=tiles, vnni=vnni)
    gen = oracle.TensorInit("normal", BF16, seed)

    def t(a):
        return torch.from_numpy(a.view(np.int16))

    replays, wants = [], []
    for _ in range(n_sets):
        Ws = [gen.fill(c, k) for c, k in zip(layers[:-1], layers[1:])]
        bs = [gen.fill(k) for k in layers[1:]]
        x = gen.fill(batch, layers[0])
        wp = [harness.pack_weight(t(W), bk, bc) for W in Ws]
        if vnni:
            wp = [harness.vnni_pack_weight(w, vnni_factor) for w in wp]
        acts = [harness.pack_activation(t(x), bn, bc).cuda()] + [torch.zeros(batch * k, dtype=torch.int16).cuda()
                                                                   for k in layers[1:]]
        replays.append(harness.MlpReplay(cfg, [w.cuda() for w in wp], [t(b).cuda() for b in bs], acts))
        ref = x
        for W, b in zip(Ws, bs):
            y = np.zeros((batch, W.shape[1]), np.uint16)
            oracle.fused_brgemm(BF16, batch, W.shape[1], W.shape[0], W.shape[0], W.shape[1], W.shape[1], 0, 0, 4, 0, 5, 4,
                                1, ref, W, y, b, 1)
            ref = y
        wants.append(ref)
    return cfg, replays, wants


def _blocked_out(cfg, replay):
    from tpp_mlir_b200 import harness

    bn, bk, _ = cfg.tiles
    return harness.unpack_activation(replay.acts[-1].reshape(cfg.batch // bn, cfg.layers[-1] // bk, bn, bk)).cpu().numpy().view(
        np.uint16)


@pytest.fixture
def pair_kernel_only():
    """Lone block-packed / VNNI-2 chains normally run on the pass kernel's GEN instantiations (mlp_chain_ft.cu);
    TPP_XSMM_CHAIN_FTG=0 keeps them on the pair-per-chain kernel (column-split items), which these tests then cover."""
    old = os.environ.get("TPP_XSMM_CHAIN_FTG")
    os.environ["TPP_XSMM_CHAIN_FTG"] = "0"
    yield
    if old is None:
        del os.environ["TPP_XSMM_CHAIN_FTG"]
    else:
        os.environ["TPP_XSMM_CHAIN_FTG"] = old


@pytest.mark.parametrize("tiles", [(32, 32, 32), (64, 64, 64), (32, 64, 64), (128, 128, 128), (256, 64, 64), (64, 256, 256)])
@pytest.mark.parametrize("vnni", [False, True])
def test_captured_tile_invokes_are_regrouped_onto_the_pair_kernel(tiles, vnni, pair_kernel_only):
    """The reference's DEFAULT call stream (benchmarks/config/omp/mlir-bf16.json:37: --tiles=32,32,32 --vnni=2; one small
    BRGEMM per (iN, iK) output block on block-packed operands) captured into a graph must not run as one launch per
    tile: the runtime folds the invokes of a layer back into one work item and the whole chain runs as ONE launch of
    the pair-per-chain kernel (4-D tensor maps over the blocked operands, VNNI-2 weights converted in the kernel)."""
    from tpp_mlir_b200 import xsmm

    cfg, (r,), (want,) = _blocked_mlp(tiles, vnni)
    n0 = xsmm.launch_count()
    with xsmm.graph_capture() as g:
        r.forward()
    name = xsmm.last_kernel()
    assert "pair256x256" in name and "_blocked" in name, name
    assert ("_vnni2" in name) == vnni, name
    g.launch()
    xsmm.sync()
    assert xsmm.launch_count() - n0 == 1, "one kernel for the whole forward pass"
    assert_close(BF16, _blocked_out(cfg, r), want)
    # replays are bit-identical
    first = _blocked_out(cfg, r).copy()
    r.acts[-1].zero_()
    g.launch()
    xsmm.sync()
    assert (_blocked_out(cfg, r) == first).all()
    g.destroy()


@pytest.mark.parametrize("tiles", [(32, 32, 32), (64, 64, 64), (32, 64, 64), (128, 128, 128), (256, 64, 64), (64, 256, 256),
                                   (256, 1024, 1024)])
@pytest.mark.parametrize("vnni", [False, True])
def test_lone_blocked_forward_runs_on_the_pass_kernel(tiles, vnni):
    """ONE forward pass of the reference's call stream (what `tpp-run -n` re-runs on one set of buffers): the layers folded
    from the tile invokes go to the GEN instantiations of the feature-major pass kernel - 128 CTAs of 32 rows x 64
    features, 5-D / 4-D tensor maps over the block-packed operands, SWIZZLE_64B sub-tiles for 32-wide k blocks, VNNI-2
    weights rewritten in shared memory - instead of a single SM pair. 3 layers; every layer's output is checked against
    the oracle, replays with poisoned intermediates are bit-identical. Flat non-VNNI weights keep the flat kernel; 32-wide
    non-VNNI weight blocks stay on the pair kernel."""
    import torch

    from tpp_mlir_b200 import harness, xsmm

    cfg, (r,), (want,) = _blocked_mlp(tiles, vnni, layers=(1024, 1024, 1024, 1024))
    n0 = xsmm.launch_count()
    with xsmm.graph_capture() as g:
        r.forward()
    name = xsmm.last_kernel()
    bn, bk, bc = tiles
    if bk == 32 and not vnni:
        assert "pair256x256_blocked" in name, name
    else:
        assert "ft64x32_fullk" in name, name
        assert ("_blocked" in name) == (tiles != (256, 1024, 1024)), name
    assert ("_vnni2" in name) == vnni, name
    g.launch()
    xsmm.sync()
    assert xsmm.launch_count() - n0 == 1, "one kernel for the whole forward pass"
    assert_close(BF16, _blocked_out(cfg, r), want)
    first = [a.clone() for a in r.acts[1:]]
    for rep in range(3):
        for a in r.acts[1:]:
            a.fill_(0x7FC0)
        g.launch()
        xsmm.sync()
        for a, f in zip(r.acts[1:], first):
            assert torch.equal(a, f), "a layer ran ahead of its input"
    g.destroy()


@pytest.mark.parametrize("tiles", [(32, 32, 32), (64, 64, 64), (256, 1024, 1024)])
@pytest.mark.parametrize("n_sets", [1, 14])
def test_vnni4_chains_run_on_the_fused_kernels(tiles, n_sets, monkeypatch):
    """mlir-gen --vnni=4 (benchmarks/config/omp/mlir-bf16.json:65-125): weights packed [k/4][n][4], the factor being what
    libxsmm_cpuid_dot_pack_factor answers (TPP_XSMM_VNNI=4). A captured chain reads flat copies of the weights that one
    small kernel in front of the chain kernel makes (vnni_flat.cu): a lone forward pass on the pass kernel, many on the
    pair-per-chain kernel - not one generic launch per tile."""
    from tpp_mlir_b200 import xsmm

    monkeypatch.setenv("TPP_XSMM_VNNI", "4")
    cfg, replays, wants = _blocked_mlp(tiles, True, layers=(1024, 1024, 1024, 1024), n_sets=n_sets, seed=5 + n_sets, vnni_factor=4)
    n0 = xsmm.launch_count()
    with xsmm.graph_capture() as g:
        for r in replays:
            r.forward()
    name = xsmm.last_kernel()
    assert name.endswith("_vnni4") or "_vnni4_" in name, name
    assert ("ft64x32_fullk" in name) if n_sets == 1 else ("pair256x256" in name), name
    for rep in range(2):
        for r in replays:
            r.acts[-1].zero_()
        g.launch()
        xsmm.sync()
        for r, want in zip(replays, wants):
            assert_close(BF16, _blocked_out(cfg, r), want)
    assert xsmm.launch_count() - n0 == 4, "two replays of (weight copy + chain kernel)"
    g.destroy()


@pytest.mark.parametrize("tiles,vnni", [((32, 32, 32), True), ((64, 64, 64), False)])
def test_few_blocked_chains_share_one_pass_kernel_launch(tiles, vnni):
    """Three operand sets of the block-packed stream in one graph: one launch of the pass kernel, the three chains
    interleaved layer by layer; seven sets: the pair-per-chain kernel (column-split items)."""
    from tpp_mlir_b200 import xsmm

    for n_sets, kernel in ((3, "3x3layers_ft64x32_fullk_blocked"), (7, "7x3layers_pair256x256_blocked")):
        cfg, replays, wants = _blocked_mlp(tiles, vnni, layers=(1024, 1024, 1024, 1024), n_sets=n_sets, seed=77 + n_sets)
        n0 = xsmm.launch_count()
        with xsmm.graph_capture() as g:
            for r in replays:
                r.forward()
        assert kernel in xsmm.last_kernel(), xsmm.last_kernel()
        g.launch()
        xsmm.sync()
        assert xsmm.launch_count() - n0 == 1
        for r, want in zip(replays, wants):
            assert_close(BF16, _blocked_out(cfg, r), want)
        g.destroy()


@pytest.mark.parametrize("tiles,vnni", [((32, 32, 32), True), ((64, 64, 64), True), ((64, 64, 64), False)])
def test_unrolled_blocked_loop_is_one_sequential_launch(tiles, vnni):
    """The benchmark loop of tpp-run unrolled 4x before capture (patches/0005), on the reference's own operands: the exact
    repeats of the chain run as one launch of the pass kernel (a plain sequence of dependent passes); VNNI-2 weights are
    un-interleaved once per graph launch into a graph-owned flat copy by a small kernel in front of it (2 launches)."""
    import torch

    from tpp_mlir_b200 import harness, xsmm

    cfg, (r,), (want,) = _blocked_mlp(tiles, vnni, layers=(1024, 1024, 1024, 1024))
    stream = torch.cuda.current_stream()
    xsmm.set_stream(stream.cuda_stream)
    loop = harness.NativeMlpLoop(cfg, r.handles, [(r.acts, r.weights, r.biases)])
    n0 = xsmm.launch_count()
    loop.run_graph_unrolled(8, 4)
    xsmm.sync()
    assert "4x3layers_ft64x32_fullk_blocked" in xsmm.last_kernel() and xsmm.last_kernel().endswith("_seq"), xsmm.last_kernel()
    assert xsmm.launch_count() - n0 == 2 * (2 if vnni else 1), "two replays of (weight copy +) one chain kernel"
    assert_close(BF16, _blocked_out(cfg, r), want)
    xsmm.set_stream(0)


@pytest.mark.parametrize("vnni", [False, True])
def test_reference_default_stream_many_sets_one_launch(vnni):
    """14 operand sets of the --tiles=32,32,32 stream (3 layers each, 768 invokes per set) captured in one graph: one
    launch, every set checked against the oracle."""
    from tpp_mlir_b200 import xsmm

    cfg, replays, wants = _blocked_mlp((32, 32, 32), vnni, layers=(1024, 1024, 1024, 1024), n_sets=14)
    n0 = xsmm.launch_count()
    with xsmm.graph_capture() as g:
        for r in replays:
            r.forward()
    name = xsmm.last_kernel()
    assert name.startswith("mlp_chain_bf16_14x3layers_pair256x256_blocked"), name
    g.launch()
    xsmm.sync()
    assert xsmm.launch_count() - n0 == 1
    for r, want in zip(replays, wants):
        assert_close(BF16, _blocked_out(cfg, r), want)
    g.destroy()


def test_shared_vnni2_weights_are_converted_once_per_launch():
    """A batch of 2048 rows (8 work items per layer chain) on the reference's default layout: every item would rewrite the
    same VNNI-2 weights in shared memory, so the launch takes one flat copy of them instead (vnni_flat.cu) and the items
    run the flat-weight instantiation. Same answer as the oracle; 2 launches per replay (copy + chain kernel)."""
    from tpp_mlir_b200 import xsmm

    cfg, replays, wants = _blocked_mlp((32, 32, 32), True, batch=2048, layers=(1024, 1024, 1024, 1024), n_sets=2, seed=61)
    with xsmm.graph_capture() as g:
        for r in replays:
            r.forward()
    assert "16x3layers_pair256x256_blocked_vnni2" in xsmm.last_kernel(), xsmm.last_kernel()
    n0 = xsmm.launch_count()
    g.launch()
    xsmm.sync()
    assert xsmm.launch_count() - n0 == 2, "weight copy + chain kernel"
    for r, want in zip(replays, wants):
        assert_close(BF16, _blocked_out(cfg, r), want)
    g.destroy()


@pytest.mark.parametrize("n_sets", [3, 8])
def test_marked_temporaries_are_ignored_where_rows_are_shared(n_sets):
    """Marks on launches that cannot honour them: 3 block-packed chains run on the pass kernel (which never discards),
    8 on the pair kernel with column-split items (other pairs read a pair's rows: no discard). Results as unmarked."""
    from tpp_mlir_b200 import xsmm

    cfg, replays, wants = _blocked_mlp((64, 64, 64), False, layers=(1024, 1024, 1024, 1024), n_sets=n_sets, seed=41)
    for r in replays:
        for a in r.acts[1:-1]:
            xsmm.mark_temporary(a)
    try:
        with xsmm.graph_capture() as g:
            for r in replays:
                r.forward()
        name = xsmm.last_kernel()
        assert ("ft64x32" in name) if n_sets == 3 else ("_split" in name), name
        for rep in range(2):
            g.launch()
            xsmm.sync()
            for r, want in zip(replays, wants):
                assert_close(BF16, _blocked_out(cfg, r), want)
        g.destroy()
    finally:
        for r in replays:
            for a in r.acts[1:-1]:
                xsmm.unmark_temporary(a)


@pytest.mark.parametrize("tiles,vnni", [((256, 1024, 1024), False), ((32, 32, 32), True), ((64, 64, 64), False)])
def test_marked_temporaries_do_not_change_results(tiles, vnni):
    """xsmm_cuda_mark_temporary on the intermediate activations (function-local buffers of the reference's generated
    kernel): the pair-per-chain kernel drops them from L2 after the next layer has consumed them. The final outputs must be
    bit-identical to the unmarked run and match the oracle; replays too (a discard issued before the last read would
    show up here)."""
    from tpp_mlir_b200 import xsmm

    n_sets = 14
    cfg, replays, wants = _blocked_mlp(tiles, vnni, layers=(1024, 1024, 1024, 1024), n_sets=n_sets, seed=31)
    with xsmm.graph_capture() as g0:
        for r in replays:
            r.forward()
    assert "pair256x256" in xsmm.last_kernel(), xsmm.last_kernel()
    g0.launch()
    xsmm.sync()
    plain = [_blocked_out(cfg, r).copy() for r in replays]
    for r in replays:
        for a in r.acts[1:-1]:
            xsmm.mark_temporary(a)
    try:
        with xsmm.graph_capture() as g1:
            for r in replays:
                r.forward()
        for rep in range(3):
            for r in replays:
                for a in r.acts[1:]:
                    a.fill_(0x7FC0)
            g1.launch()
            xsmm.sync()
            for r, want, p in zip(replays, wants, plain):
                got = _blocked_out(cfg, r)
                assert np.array_equal(got, p), "marked temporaries changed the result"
                assert_close(BF16, got, want)
        g1.destroy()
    finally:
        for r in replays:
            for a in r.acts[1:-1]:
                xsmm.unmark_temporary(a)
    g0.destroy()


@pytest.mark.parametrize("dtype,batch,layers,tiles", [(F32, 256, (1024, 1024, 1024), (32, 32, 32)),
                                                      (F32, 128, (512, 352), (32, 32, 32)),
                                                      (BF16, 128, (1024, 2304), (64, 48, 64)),
                                                      (BF16, 96, (256, 512, 256), (32, 32, 32))])
def test_grids_no_fused_kernel_takes_run_as_one_generic_launch_per_layer(dtype, batch, layers, tiles):
    """Tile invokes the tcgen05 chain kernels do not take - f32 (the reference's fp32 MLP configs,
    benchmarks/config/base/base.json), a batch that is no multiple of 256 rows (benchmarks/config/fc: --batch=128) - are
    still folded into layers under capture and go out as ONE launch of the generic kernel per layer (one z-slice per tile)
    instead of one launch per tile; so are 48-wide tiles (benchmarks/config/fc: --tiles=64,48,64) and batches that are no
    multiple of 128 rows. Checked against the oracle (f32: f64-accumulate mode, 1e-5)."""
    import torch

    from tpp_mlir_b200 import harness, xsmm

    bn, bk, bc = tiles
    cfg = harness.MlpConfig(batch=batch, layers=layers, tiles=tiles, dtype=dtype)
    gen = oracle.TensorInit("normal", dtype, 77)
    Ws = [gen.fill(c, k) for c, k in zip(layers[:-1], layers[1:])]
    bs = [gen.fill(k) for k in layers[1:]]
    x = gen.fill(batch, layers[0])
    tdt = torch.float32 if dtype == F32 else torch.int16

    def t(a):
        return torch.from_numpy(a if dtype == F32 else a.view(np.int16))

    wp = [harness.pack_weight(t(W), bk, bc).cuda() for W in Ws]
    acts = [harness.pack_activation(t(x), bn, bc).cuda()] + [torch.zeros(batch * k, dtype=tdt).cuda() for k in layers[1:]]
    r = harness.MlpReplay(cfg, wp, [t(b).cuda() for b in bs], acts)
    n0 = xsmm.launch_count()
    with xsmm.graph_capture() as g:
        r.forward()
    assert "_grid" in xsmm.last_kernel() and "simt" in xsmm.last_kernel(), xsmm.last_kernel()
    g.launch()
    xsmm.sync()
    assert xsmm.launch_count() - n0 == len(layers) - 1, "one launch per layer"
    oracle.set_acc_mode(1 if dtype == F32 else 0)
    try:
        ref = x
        for W, b in zip(Ws, bs):
            y = np.zeros((batch, W.shape[1]), np.float32 if dtype == F32 else np.uint16)
            oracle.fused_brgemm(dtype, batch, W.shape[1], W.shape[0], W.shape[0], W.shape[1], W.shape[1], 0, 0, 4, 0, 5, 4, 1, ref, W,
                                y, b, 1)
            ref = y
    finally:
        oracle.set_acc_mode(0)
    got = harness.unpack_activation(acts[-1].reshape(batch // bn, layers[-1] // bk, bn, bk)).cpu().numpy()
    got = got if dtype == F32 else got.view(np.uint16)
    assert_close(dtype, got, ref)
    g.destroy()


@pytest.mark.parametrize("batch,layers,tiles,vnni", [(128, (1024, 4096), (64, 64, 64), False), (128, (768, 768), (32, 32, 32), True),
                                                      (384, (1024, 1024, 1024), (64, 64, 64), False),
                                                      (128, (768, 3072), (128, 256, 64), False)])
def test_batches_of_128_rows_run_on_the_pair_kernel(batch, layers, tiles, vnni):
    """benchmarks/config/fc and matmul run --batch=128 (and any batch that is an odd number of 128-row halves): the last
    work item of the pair-per-chain kernel then has 128 rows only; its second CTA works on rows beyond the operands, which
    TMA zero-fills on loads and drops on stores. One launch, the oracle's answer, nothing written outside the output."""
    import torch

    from tpp_mlir_b200 import harness, xsmm

    bn, bk, bc = tiles
    cfg = harness.MlpConfig(batch=batch, layers=layers, tiles=tiles, vnni=vnni)
    gen = oracle.TensorInit("normal", BF16, 99)
    Ws = [gen.fill(c, k) for c, k in zip(layers[:-1], layers[1:])]
    bs = [gen.fill(k) for k in layers[1:]]
    x = gen.fill(batch, layers[0])

    def t(a):
        return torch.from_numpy(a.view(np.int16))

    wp = [harness.pack_weight(t(W), bk, bc) for W in Ws]
    if vnni:
        wp = [harness.vnni_pack_weight(w) for w in wp]
    guard = 4096
    acts = [harness.pack_activation(t(x), bn, bc).cuda()]
    raw = []
    for k in layers[1:]:   # outputs with a guard band behind them
        buf = torch.full((batch * k + guard,), 0x1234, dtype=torch.int16, device="cuda")
        raw.append(buf)
        acts.append(buf[:batch * k])
    r = harness.MlpReplay(cfg, [w.cuda() for w in wp], [t(b).cuda() for b in bs], acts)
    n0 = xsmm.launch_count()
    with xsmm.graph_capture() as g:
        r.forward()
    assert "pair256x256" in xsmm.last_kernel() or "ft64x32" in xsmm.last_kernel(), xsmm.last_kernel()
    g.launch()
    xsmm.sync()
    assert xsmm.launch_count() - n0 == 1
    ref = x
    for W, b in zip(Ws, bs):
        y = np.zeros((batch, W.shape[1]), np.uint16)
        oracle.fused_brgemm(BF16, batch, W.shape[1], W.shape[0], W.shape[0], W.shape[1], W.shape[1], 0, 0, 4, 0, 5, 4, 1, ref, W, y, b, 1)
        ref = y
    got = harness.unpack_activation(acts[-1].reshape(batch // bn, layers[-1] // bk, bn, bk)).cpu().numpy().view(np.uint16)
    assert_close(BF16, got, ref)
    for buf, k in zip(raw, layers[1:]):
        assert (buf[batch * k:] == 0x1234).all(), "rows beyond the batch were written"
    g.destroy()


def test_regrouping_keeps_irregular_invoke_streams_as_they_are():
    """Tile invokes that do NOT walk a regular grid (here: the column blocks of a layer visited in a shuffled order) are
    launched one by one, in program order - same answer, no fused kernel."""
    import torch

    from tpp_mlir_b200 import harness, xsmm

    cfg, (r,), (want,) = _blocked_mlp((64, 64, 64), False, layers=(256, 256))
    bn, bk, bc = cfg.tiles
    order = [(i, j) for i in range(cfg.batch // bn) for j in (2, 0, 3, 1)]
    with xsmm.graph_capture() as g:
        for i_n, i_k in order:
            xsmm.fused_brgemm_invoke(cfg.dtype, r.handles[0], r.acts[0], i_n * (256 // bc) * bn * bc, r.weights[0],
                                     i_k * (256 // bc) * bc * bk, r.acts[1], (i_n * (256 // bk) + i_k) * bn * bk, r.biases[0],
                                     i_k * bk, 256 // bc)
    assert "pair" not in xsmm.last_kernel(), xsmm.last_kernel()
    g.launch()
    xsmm.sync()
    assert_close(BF16, _blocked_out(cfg, r), want)
    g.destroy()


def test_single_blocked_layer_runs_on_the_pair_kernel():
    """One layer (no chain) as a grid of tile invokes under capture: still one tensor-core launch."""
    from tpp_mlir_b200 import xsmm

    cfg, (r,), (want,) = _blocked_mlp((32, 32, 32), True, layers=(512, 768))
    with xsmm.graph_capture() as g:
        r.forward()
    assert "pair256x256_blocked_vnni2" in xsmm.last_kernel(), xsmm.last_kernel()
    g.launch()
    xsmm.sync()
    assert_close(BF16, _blocked_out(cfg, r), want)
    g.destroy()


@pytest.mark.parametrize("tiles,vnni", [("32,32,32", 2), ("64,64,64", 0), ("256,256,256", 0)])
def test_tpp_run_standin_modes_agree(tiles, vnni):
    """The native tpp-run stand-in (csrc/harness/tpp_run_standin.cpp) in its three residency modes - plain host pointers,
    device arguments, captured graph - computes the same forward pass (checksums within the bf16 tolerance) and the
    graph mode runs it as one fused launch."""
    import json
    import subprocess

    from tpp_mlir_b200 import _build

    exe = _build.standin_path()
    if not os.path.exists(exe):
        exe = _build.build_standin()
    rows = {}
    for mode in ("strict", "device", "lazy", "graph"):
        r = subprocess.run([exe, "--batch", "256", "--layers", "256,512,256", "--tiles", tiles, "--vnni", str(vnni), "-n", "2",
                            "--mode", mode], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stderr
        rows[mode] = json.loads(r.stdout.strip().splitlines()[-1])
    ref = rows["strict"]["checksum"]
    assert ref > 0
    for mode in ("device", "lazy", "graph"):
        assert abs(rows[mode]["checksum"] - ref) <= 1e-2 * abs(ref), rows
    for mode in ("lazy", "graph"):
        assert "pair256x256" in rows[mode]["kernel"] or "chain" in rows[mode]["kernel"], rows[mode]
    assert rows["lazy"]["launches"] < rows["device"]["launches"], rows


@pytest.mark.parametrize("tiles", [(256, 1024, 1024), (64, 64, 64), (32, 32, 32)])
@pytest.mark.parametrize("with_zero", [False, True])
def test_unfused_layers_are_combined_under_capture(tiles, with_zero):
    """The UNFUSED form of an MLP layer (what the pipeline emits when CombineXsmmOp does not fire, SURVEY.md Appendix A:
    [unary zero] -> brgemm -> binary add(bias, bcast_col_in0, in place) -> unary relu(in place)) captured into a graph is
    combined by the runtime the way lib/TPP/Transforms/CombineXsmmPass.cpp:31-145 would have: add / relu become the
    BRGEMM's epilogue, a zero overwritten by a beta=1 BRGEMM becomes beta_0, and the two layers still chain into ONE
    launch. Checked against the oracle's fused op (one rounding) and, within the bf16 tolerance, its unfused sequence."""
    import torch

    from tpp_mlir_b200 import harness, xsmm

    bn, bk, bc = tiles
    layers = (1024, 1024, 1024) if bk == 1024 else (512, 512, 512)
    cfg, (r,), (want,) = _blocked_mlp(tiles, False, layers=layers)
    gflags = (0 if with_zero else xsmm.GEMM_FLAG_BETA_0) | 64 | 128
    hb = xsmm.brgemm_dispatch(BF16, bn, bk, bc, bc, bk, bk, bn * bc, bc * bk, gflags)
    hz = xsmm.unary_dispatch(xsmm.UNARY_ZERO, BF16, bn, bk, bk, bk, 0)
    ha = xsmm.binary_dispatch(xsmm.BINARY_ADD, BF16, bn, bk, bk, bk, bk, xsmm.BINARY_FLAG_BCAST_COL_IN_0)
    hr = xsmm.unary_dispatch(xsmm.UNARY_RELU, BF16, bn, bk, bk, bk, 0)

    def forward():
        for l in range(2):
            c, k = layers[l], layers[l + 1]
            nb_c, nb_k = c // bc, k // bk
            for i_n in range(cfg.batch // bn):
                for i_k in range(nb_k):
                    off_c = (i_n * nb_k + i_k) * bn * bk
                    if with_zero:
                        xsmm.unary_invoke(BF16, hz, r.acts[l + 1], off_c, r.acts[l + 1], off_c)
                    xsmm.brgemm_invoke(BF16, hb, r.acts[l], i_n * nb_c * bn * bc, r.weights[l], i_k * nb_c * bc * bk,
                                       r.acts[l + 1], off_c, nb_c)
                    xsmm.binary_invoke(BF16, ha, r.biases[l], i_k * bk, r.acts[l + 1], off_c, r.acts[l + 1], off_c)
                    xsmm.unary_invoke(BF16, hr, r.acts[l + 1], off_c, r.acts[l + 1], off_c)

    n0 = xsmm.launch_count()
    with xsmm.graph_capture() as g:
        forward()
    name = xsmm.last_kernel()
    g.launch()
    xsmm.sync()
    assert xsmm.launch_count() - n0 == 1, (xsmm.launch_count() - n0, name)
    assert "mlp_chain" in name, name
    fused = _blocked_out(cfg, r)
    assert_close(BF16, fused, want)
    g.destroy()
    # the same stream issued directly (no capture): every op is its own launch, three roundings per layer
    for a in r.acts[1:]:
        a.zero_()
    n0 = xsmm.launch_count()
    forward()
    xsmm.sync()
    assert xsmm.launch_count() - n0 > 2
    assert_close(BF16, _blocked_out(cfg, r), want)


def test_lazy_mode_queues_invokes_and_keeps_program_order():
    """Lazy mode (xsmm_cuda_set_lazy): the uncaptured tile-invoke stream of a 2-layer block-packed MLP is queued and goes
    out as one fused launch at sync(); ops that cannot be queued (a unary relu on another buffer, a binary op) flush the
    queue first, so program order is kept; turning lazy mode off flushes."""
    import torch

    from tpp_mlir_b200 import xsmm

    cfg, (r,), (want,) = _blocked_mlp((32, 32, 32), True)
    xsmm.set_lazy(True)
    try:
        n0 = xsmm.launch_count()
        r.forward()
        assert xsmm.launch_count() == n0, "nothing is launched before a flush point"
        xsmm.sync()
        assert xsmm.launch_count() - n0 == 1, xsmm.launch_count() - n0
        assert "pair256x256_blocked_vnni2" in xsmm.last_kernel(), xsmm.last_kernel()
        assert_close(BF16, _blocked_out(cfg, r), want)
        # program order against a consumer that is not a BRGEMM: out2 = relu(result) must see the NEW result
        for a in r.acts[1:]:
            a.zero_()
        out2 = torch.zeros_like(r.acts[-1])
        hr = xsmm.unary_dispatch(xsmm.UNARY_RELU, BF16, 256, 1024, 1024, 1024, 0)
        r.forward()
        xsmm.unary_invoke(BF16, hr, r.acts[-1], 0, out2, 0)   # flushes the queued layers first
        xsmm.set_lazy(False)
        xsmm.sync()
        np.testing.assert_array_equal(out2.cpu().numpy(), r.acts[-1].cpu().numpy())
        assert_close(BF16, _blocked_out(cfg, r), want)
    finally:
        xsmm.set_lazy(False)


def test_perf_timer_includes_async_launches():
    import torch

    from tpp_mlir_b200 import xsmm

    A = torch.ones(16, 1024, 1024, dtype=torch.bfloat16, device="cuda")
    C = torch.zeros(1024, 1024, dtype=torch.bfloat16, device="cuda")
    h = xsmm.brgemm_dispatch(BF16, 1024, 1024, 1024, 1024, 1024, 1024, 1 << 20, 1 << 20, 4)
    xsmm.brgemm_invoke(BF16, h, A, 0, A, 0, C, 0, 16)
    n0 = xsmm.launch_count()
    t0 = xsmm.perf_start_timer()
    for _ in range(5):
        xsmm.brgemm_invoke(BF16, h, A, 0, A, 0, C, 0, 16)
    dt = xsmm.perf_stop_timer(t0)
    assert xsmm.launch_count() - n0 == 5
    # 5 x 34.4 GFLOP cannot finish faster than the nominal 2.25 PFLOP/s peak allows
    assert dt > 5 * 34.36e9 / 2.25e15
    assert float(C[0, 0]) == 16384.0


def test_dispatch_is_cached_and_validates():
    import subprocess
    import sys

    from tpp_mlir_b200 import xsmm

    h1 = xsmm.brgemm_dispatch(BF16, 64, 64, 64, 64, 64, 64, 4096, 4096, 4)
    h2 = xsmm.brgemm_dispatch(BF16, 64, 64, 64, 64, 64, 64, 4096, 4096, 4)
    h3 = xsmm.brgemm_dispatch(BF16, 64, 64, 64, 64, 64, 64, 4096, 4096, 0)
    assert h1 == h2 and h1 != h3
    assert xsmm.handle_kernel(h1).startswith("brgemm_tc_bf16")
    # lda < k violates the op verifier: the reference's dispatch prints and exit(-1)s; so does this one
    code = "from tpp_mlir_b200 import xsmm; xsmm.brgemm_dispatch(2, 64, 64, 64, 32, 64, 64, 0, 0, 4)"
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert p.returncode != 0 and "failed to generate brgemm func" in p.stderr


def test_graph_capture_replays_the_invoke_sequence():
    """xsmm_cuda_graph_begin/end: the captured 2-layer invoke chain replays bit-identically."""
    import torch

    from tpp_mlir_b200 import xsmm

    gen = oracle.TensorInit("normal", BF16, 7)

    def dev_t(a):
        return torch.from_numpy(a.view(np.int16)).cuda()

    x, W1, b1, W2, b2 = (dev_t(gen.fill(*s)) for s in ((256, 512), (512, 512), (512,), (512, 512), (512,)))
    y1 = torch.zeros(256, 512, dtype=torch.int16, device="cuda")
    y2 = torch.zeros(256, 512, dtype=torch.int16, device="cuda")
    h = xsmm.fused_brgemm_dispatch(BF16, 256, 512, 512, 512, 512, 512, 0, 0, 4, 0, 5, 4, 1)

    def chain():
        xsmm.fused_brgemm_invoke(BF16, h, x, 0, W1, 0, y1, 0, b1, 0, 1)
        xsmm.fused_brgemm_invoke(BF16, h, y1, 0, W2, 0, y2, 0, b2, 0, 1)

    chain()
    xsmm.sync()
    want = y2.clone()
    n0 = xsmm.launch_count()
    with xsmm.graph_capture() as g:
        chain()
    assert xsmm.launch_count() == n0  # capture records, it does not execute
    for _ in range(3):
        y1.zero_()
        y2.zero_()
        g.launch()
        xsmm.sync()
        assert torch.equal(y2, want)
    assert xsmm.launch_count() == n0 + 6
    g.destroy()


def test_concurrent_invokes_from_many_threads(orc):
    """The reference runs invokes concurrently from OpenMP threads on disjoint C tiles (--def-parallel,
    lib/TPP/DefaultPipeline.cpp:179-180): dispatch cache, per-thread streams / staging / tensor-map caches must be
    thread-safe. 8 threads, each with its own stream, mix of tensor-core, generic and eltwise kernels and of
    device / plain-host operands."""
    import threading

    import torch

    from tpp_mlir_b200 import xsmm

    errors = []

    def worker(tid):
        try:
            rng = np.random.default_rng(100 + tid)
            stream = torch.cuda.Stream()
            be = AbiBackend("device" if tid % 2 == 0 else "host")
            with torch.cuda.stream(stream):
                for it in range(6):
                    m, n, k, batch = 64 + 32 * (tid % 3), 128, 64 * (1 + it % 2), 1 + it % 3
                    A, B = rnd(rng, BF16, (batch, m, k)), rnd(rng, BF16, (batch, k, n))
                    bias, C0 = rnd(rng, BF16, (n,)), rnd(rng, BF16, (m, n))
                    g, o = C0.copy(), C0.copy()
                    be.fused_brgemm(BF16, m, n, k, k, n, n, m * k, k * n, 0, 0, 5, 4, 1, A, 0, B, 0, g, 0, bias, 0, batch)
                    orc.fused_brgemm(BF16, m, n, k, k, n, n, m * k, k * n, 0, 0, 5, 4, 1, A, 0, B, 0, o, 0, bias, 0, batch)
                    assert_close(BF16, g, o)
                    x = rnd(rng, F32, (33, 47))
                    yg, yo = np.zeros((47, 33), np.float32), np.zeros((47, 33), np.float32)
                    be.unary(29, F32, 33, 47, 47, 33, 0, x, 0, yg, 0)
                    orc.unary(29, F32, 33, 47, 47, 33, 0, x, 0, yo, 0)
                    np.testing.assert_array_equal(yg, yo)
        except Exception as e:  # noqa: BLE001
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    xsmm.sync()
    assert not errors, errors


@pytest.mark.parametrize("layers,view", [(3, "flat"), (2, "flat"), (4, "flat"), (3, "k64xb16")])
def test_captured_mlp_chain_is_fused_and_matches(layers, view):
    """Graph capture of consecutive layers (C of one = A of the next) launches ONE persistent chain kernel
    (SURVEY 8f-2). Every layer's output is within tolerance of the oracle AND within one bf16 rounding step of the
    per-layer kernels (the chain accumulates the full reduction in one TMEM tile, the stand-alone kernel adds four
    split-K partials: the f32 sums may differ in the last bit before the single bf16 rounding); replays are
    bit-identical to each other."""
    import torch

    from tpp_mlir_b200 import xsmm

    gen = oracle.TensorInit("normal", BF16, 11 + layers)
    Ws = [gen.fill(1024, 1024) for _ in range(layers)]
    bs = [gen.fill(1024) for _ in range(layers)]
    x = gen.fill(256, 1024)

    def dev_t(a):
        return torch.from_numpy(a.view(np.int16)).cuda()

    dW, db = [dev_t(w) for w in Ws], [dev_t(b) for b in bs]
    acts = [dev_t(x)] + [torch.zeros(256, 1024, dtype=torch.int16, device="cuda") for _ in range(layers)]
    if view == "flat":
        h = xsmm.fused_brgemm_dispatch(BF16, 256, 1024, 1024, 1024, 1024, 1024, 256 * 1024, 1 << 20, 4 | 64 | 128, 0, 5, 4, 1)
        nb = 1
    else:  # strided view of the same math: k = 64 x batch 16
        h = xsmm.fused_brgemm_dispatch(BF16, 256, 1024, 64, 1024, 1024, 1024, 64, 64 * 1024, 4 | 64 | 128, 0, 5, 4, 1)
        nb = 16

    def forward():
        for l in range(layers):
            xsmm.fused_brgemm_invoke(BF16, h, acts[l], 0, dW[l], 0, acts[l + 1], 0, db[l], 0, nb)

    forward()
    xsmm.sync()
    direct = [a.clone() for a in acts[1:]]
    n0 = xsmm.launch_count()
    with xsmm.graph_capture() as g:
        forward()
    for a in acts[1:]:
        a.zero_()
    g.launch()
    xsmm.sync()
    assert xsmm.launch_count() - n0 == 1, "the captured chain must be ONE kernel launch"
    assert xsmm.last_kernel().startswith(f"mlp_chain_bf16_{layers}layers"), xsmm.last_kernel()
    for got, want in zip(acts[1:], direct):
        gi, wi = got.cpu().numpy().view(np.uint16).astype(np.int32), want.cpu().numpy().view(np.uint16).astype(np.int32)
        assert np.abs(gi - wi).max() <= 1, "chain vs per-layer kernels: more than one bf16 ulp apart"
        assert (gi != wi).mean() < 1e-3
    first = [a.clone() for a in acts[1:]]
    ref = x
    for W, b in zip(Ws, bs):
        y = np.zeros((256, 1024), np.uint16)
        oracle.fused_brgemm(BF16, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4, 0, 5, 4, 1, ref, W, y, b, 1)
        ref = y
    assert_close(BF16, acts[-1].cpu().numpy().view(np.uint16), ref)
    # replay again: the grid-barrier counters are monotonic across launches
    for a in acts[1:]:
        a.fill_(0x7FC0)  # poison (bf16 NaN): a layer that ran ahead of its producers would propagate it
    g.launch()
    g.launch()
    xsmm.sync()
    for a, f in zip(acts[1:], first):
        assert torch.equal(a, f)
    g.destroy()


def _mlp_chain_setup(n_chains, layers, seed):
    import torch

    gen = oracle.TensorInit("normal", BF16, seed)

    def dev_t(a):
        return torch.from_numpy(a.view(np.int16)).cuda()

    chains = []
    for _ in range(n_chains):
        Ws = [dev_t(gen.fill(1024, 1024)) for _ in range(layers)]
        bs = [dev_t(gen.fill(1024)) for _ in range(layers)]
        acts = [dev_t(gen.fill(256, 1024))] + [torch.zeros(256, 1024, dtype=torch.int16, device="cuda") for _ in range(layers)]
        chains.append((acts, Ws, bs))
    return chains


def _ulp_diff(a, b):
    return np.abs(a.cpu().numpy().view(np.uint16).astype(np.int32) - b.cpu().numpy().view(np.uint16).astype(np.int32)).max()


def test_captured_independent_chains_share_one_interleaved_launch():
    """Several independent forward passes captured in one graph (the benchmark's rotating operand sets) become ONE
    launch of the chain kernel with the chains interleaved three at a time; every layer of every chain stays within one
    bf16 rounding step of the per-layer kernels, replays are bit-identical (poisoned intermediates)."""
    import torch

    from tpp_mlir_b200 import xsmm

    layers, n_chains = 3, 7
    chains = _mlp_chain_setup(n_chains, layers, 29)
    h = xsmm.fused_brgemm_dispatch(BF16, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4 | 64 | 128, 0, 5, 4, 1)

    def forward(c):
        acts, Ws, bs = c
        for l in range(layers):
            xsmm.fused_brgemm_invoke(BF16, h, acts[l], 0, Ws[l], 0, acts[l + 1], 0, bs[l], 0, 1)

    for c in chains:
        forward(c)
    xsmm.sync()
    direct = [[a.clone() for a in c[0][1:]] for c in chains]
    with xsmm.graph_capture() as g:
        for c in chains:
            forward(c)
    assert xsmm.last_kernel().startswith(f"mlp_chain_bf16_{n_chains}x{layers}layers"), xsmm.last_kernel()
    first = None
    for rep in range(3):
        for c in chains:
            for a in c[0][1:]:
                a.fill_(0x7FC0)
        n0 = xsmm.launch_count()
        g.launch()
        xsmm.sync()
        assert xsmm.launch_count() - n0 == 1, "independent chains must share one launch"
        got = [[a.clone() for a in c[0][1:]] for c in chains]
        if first is None:
            first = got
            for gc, dc in zip(got, direct):
                for a, d in zip(gc, dc):
                    assert _ulp_diff(a, d) <= 1
        else:
            for gc, fc in zip(got, first):
                for a, f in zip(gc, fc):
                    assert torch.equal(a, f)
    g.destroy()


def test_captured_dependent_or_aliased_chains_are_not_interleaved():
    """A chain that reads another chain's output, or writes buffers another chain uses, must not run interleaved with
    it: such chains get their own launches (stream order), and the results match running them one after the other."""
    import torch

    from tpp_mlir_b200 import xsmm

    layers = 2
    (a1, W1, b1), (a2, W2, b2) = _mlp_chain_setup(2, layers, 31)
    h = xsmm.fused_brgemm_dispatch(BF16, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4 | 64 | 128, 0, 5, 4, 1)
    a2[0] = a1[-1]   # chain 2 consumes chain 1's output ...
    extra = torch.zeros(256, 1024, dtype=torch.int16, device="cuda")

    def forward():
        for l in range(layers):
            xsmm.fused_brgemm_invoke(BF16, h, a1[l], 0, W1[l], 0, a1[l + 1], 0, b1[l], 0, 1)
        # ... through an unrelated op in between, so that the two chains stay two chains
        hz = xsmm.unary_dispatch(2, BF16, 256, 1024, 1024, 1024, 0)
        xsmm.unary_invoke(BF16, hz, extra, 0, extra, 0)
        for l in range(layers):
            xsmm.fused_brgemm_invoke(BF16, h, a2[l], 0, W2[l], 0, a2[l + 1], 0, b2[l], 0, 1)
        # and a third pass over chain 1's buffers again (aliases chain 1 completely)
        for l in range(layers):
            xsmm.fused_brgemm_invoke(BF16, h, a1[l], 0, W1[l], 0, a1[l + 1], 0, b1[l], 0, 1)

    forward()
    xsmm.sync()
    want = [t.clone() for t in a1[1:] + a2[1:]]
    with xsmm.graph_capture() as g:
        forward()
    for t in a1[1:] + a2[1:]:
        t.fill_(0x7FC0)
    n0 = xsmm.launch_count()
    g.launch()
    xsmm.sync()
    assert xsmm.launch_count() - n0 == 4, "three chain launches + the zero kernel"
    for t, w in zip(a1[1:] + a2[1:], want):
        assert _ulp_diff(t, w) <= 1
    g.destroy()


def test_chain_with_weights_written_just_before():
    """Weights / bias produced by a kernel issued right before the chain (here: identity copies through the ABI) must
    not be fetched before the programmatic-launch wait; the chain still runs fused and matches."""
    import torch

    from tpp_mlir_b200 import xsmm

    layers = 3
    ((acts, Ws, bs),) = _mlp_chain_setup(1, layers, 37)
    h = xsmm.fused_brgemm_dispatch(BF16, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4 | 64 | 128, 0, 5, 4, 1)
    hc = xsmm.unary_dispatch(1, BF16, 1024, 1024, 1024, 1024, 0)
    W_live = [torch.zeros_like(w) for w in Ws]

    def forward():
        for l in range(layers):
            xsmm.unary_invoke(BF16, hc, Ws[l], 0, W_live[l], 0)     # "training step": fresh weights every forward
        for l in range(layers):
            xsmm.fused_brgemm_invoke(BF16, h, acts[l], 0, W_live[l], 0, acts[l + 1], 0, bs[l], 0, 1)

    forward()
    xsmm.sync()
    want = [a.clone() for a in acts[1:]]
    with xsmm.graph_capture() as g:
        forward()
    assert xsmm.last_kernel().startswith("mlp_chain_bf16_3layers"), xsmm.last_kernel()
    for rep in range(3):
        for w in W_live:
            w.zero_()
        for a in acts[1:]:
            a.fill_(0x7FC0)
        g.launch()
    xsmm.sync()
    for a, w in zip(acts[1:], want):
        assert _ulp_diff(a, w) <= 1
    g.destroy()


def test_split_k_chain_kernel_still_matches():
    """The first chain design (4-CTA split-K clusters + grid barrier, TPP_XSMM_CHAIN=s) stays available and
    bit-identical to the per-layer kernels; the option is read once per process, hence the subprocess."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TPP_XSMM_CHAIN="s")
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "chain_check.py"), "3", "2"], env=env, cwd=root,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "splitk4" in out.stdout and "differing_elems=0" in out.stdout, out.stdout
    assert "replay determinism: 0 unstable" in out.stdout, out.stdout


def test_captured_non_chain_sequences_are_not_fused():
    import torch

    from tpp_mlir_b200 import xsmm

    a = (torch.rand(256, 1024, device="cuda") * 0.1).bfloat16()
    w = (torch.rand(1024, 1024, device="cuda") * 0.1).bfloat16()
    c1 = torch.zeros(256, 1024, device="cuda", dtype=torch.bfloat16)
    c2 = torch.zeros(256, 1024, device="cuda", dtype=torch.bfloat16)
    h = xsmm.brgemm_dispatch(BF16, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4)
    n0 = xsmm.launch_count()
    with xsmm.graph_capture() as g:   # two independent GEMMs on the same input: not a chain
        xsmm.brgemm_invoke(BF16, h, a, 0, w, 0, c1, 0, 1)
        xsmm.brgemm_invoke(BF16, h, a, 0, w, 0, c2, 0, 1)
    g.launch()
    xsmm.sync()
    assert xsmm.launch_count() - n0 == 2
    assert torch.equal(c1, c2)
    ref = (a.float() @ w.float())
    assert ((c1.float() - ref).abs().max() / ref.abs().max()).item() < 1e-2
    g.destroy()


@pytest.mark.parametrize("mode", ["async", "grouped", "streams"])
def test_pipelined_steps_keep_apart(mode):
    """Three steps in flight, each with its own host buffers: every slot's output must be the oracle's answer for
    THAT slot's input, every time. async = upload_async / download_async / wait_host on the library's copy streams;
    grouped = three steps in one captured graph (copies as parallel branches); streams = one stream and one captured
    step graph per slot (xsmm_cuda_stream_create; per-stream kernel scratch keeps overlapping chain kernels apart)."""
    import torch

    from tpp_mlir_b200 import harness, xsmm

    cfg = harness.MlpConfig(batch=256, layers=(1024, 1024, 1024, 1024), tiles=(256, 1024, 1024), dtype=BF16)
    gen = oracle.TensorInit("normal", BF16, 321)
    Ws = [gen.fill(1024, 1024) for _ in range(3)]
    bs = [gen.fill(1024) for _ in range(3)]
    xs = [gen.fill(256, 1024) for _ in range(3)]

    def pinned(a):
        return torch.from_numpy(a.view(np.int16).copy()).contiguous().pin_memory()

    h_w, h_b = [pinned(w) for w in Ws], [pinned(b) for b in bs]
    slots = [[pinned(x)] + [torch.zeros(256 * 1024, dtype=torch.int16).pin_memory() for _ in range(3)] for x in xs]
    regs = h_w + h_b + [t for a in slots for t in a]
    for t in regs:
        xsmm.register_host(t, upload=True)
    try:
        h = xsmm.fused_brgemm_dispatch(BF16, 256, 1024, 1024, 1024, 1024, 1024, 256 * 1024, 1 << 20, 4 | 64 | 128, 0, 5,
                                       4, 1)
        loop = harness.NativeMlpLoop(cfg, [h] * 3, [(a, h_w, h_b) for a in slots])
        want = []
        for x in xs:
            ref = x
            for W, b in zip(Ws, bs):
                y = np.zeros((256, 1024), np.uint16)
                oracle.fused_brgemm(BF16, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4, 0, 5, 4, 1, ref, W, y, b, 1)
                ref = y
            want.append(ref)
        first = None
        for rounds in (3, 60, 61):
            for a in slots:
                a[-1].zero_()
            loop.run_e2e_pipelined(rounds, mode=mode)
            outs = [a[-1].numpy().view(np.uint16).reshape(256, 1024).copy() for a in slots]
            for o, w in zip(outs, want):
                assert_close(BF16, o, w)
            if first is None:
                first = outs
            else:   # bit-identical from replay to replay
                for o, f in zip(outs, first):
                    assert np.array_equal(o, f)
        assert xsmm.get_stream() in (None, 0), "the caller's stream is restored"
    finally:
        xsmm.sync()
        for t in regs:
            xsmm.unregister_host(t)


# ---- pair-per-chain kernel (mlp_chain_pair_kernel): many independent chains in one captured graph -----------------
def _is_pair_kernel(name, base):
    """`base`, optionally with the column-split tag: launches with few row blocks deal every layer's output tiles out to
    2 / 4 / 8 / 16 SM pairs per row block (mlp_chain_pair.cu: PcItem::nslices)."""
    import re

    return re.fullmatch(re.escape(base) + r"(_split(2|4|8|16))?", name) is not None


def _oracle_mlp(x, Ws, bs, m=256, relu=5, bias=True):
    ref = x
    for W, b in zip(Ws, bs):
        y = np.zeros((m, W.shape[1]), np.uint16)
        oracle.fused_brgemm(BF16, m, W.shape[1], W.shape[0], W.shape[0], W.shape[1], W.shape[1], 0, 0, 4, 0, relu,
                            4 if bias else 0, 1 if bias else 0, ref, W, y, b if bias else None, 1)
        ref = y
    return ref


@pytest.mark.parametrize("n_chains,layers,view", [(13, 3, "flat"), (20, 2, "flat"), (16, 4, "k64xb16"), (80, 3, "flat")])
def test_many_captured_chains_run_on_cta_pairs(n_chains, layers, view):
    """>= 12 independent forward passes captured in one graph become ONE launch of the pair-per-chain kernel (one CTA
    pair per chain, cta_group::2 256 x 256 tiles). Every layer of every chain is within one bf16 rounding step of the
    per-layer kernels, sampled chains match the oracle within the bf16 tolerance, replays with poisoned intermediates
    are bit-identical (a layer that ran ahead of its own output would propagate the poison). 80 chains > 74 pairs:
    pairs take a second chain."""
    import torch

    from tpp_mlir_b200 import xsmm

    gen = oracle.TensorInit("normal", BF16, 100 + n_chains)
    hW = [gen.fill(1024, 1024) for _ in range(layers)]
    hb = [gen.fill(1024) for _ in range(layers)]

    def dev_t(a):
        return torch.from_numpy(a.view(np.int16)).cuda()

    chains, xs = [], []
    for c in range(n_chains):
        x = gen.fill(256, 1024)
        xs.append(x)
        # per-chain parameters: the base matrices rolled by c rows (a chain reading another chain's table entry fails)
        Ws = [torch.roll(dev_t(w), shifts=c, dims=0).contiguous() for w in hW]
        bs = [torch.roll(dev_t(b), shifts=c, dims=0).contiguous() for b in hb]
        acts = [dev_t(x)] + [torch.zeros(256, 1024, dtype=torch.int16, device="cuda") for _ in range(layers)]
        chains.append((acts, Ws, bs))
    if view == "flat":
        h = xsmm.fused_brgemm_dispatch(BF16, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4 | 64 | 128, 0, 5, 4, 1)
        nb = 1
    else:
        h = xsmm.fused_brgemm_dispatch(BF16, 256, 1024, 64, 1024, 1024, 1024, 64, 64 * 1024, 4 | 64 | 128, 0, 5, 4, 1)
        nb = 16

    def forward(c):
        acts, Ws, bs = c
        for l in range(layers):
            xsmm.fused_brgemm_invoke(BF16, h, acts[l], 0, Ws[l], 0, acts[l + 1], 0, bs[l], 0, nb)

    for c in chains:
        forward(c)
    xsmm.sync()
    direct = [[a.clone() for a in c[0][1:]] for c in chains]
    with xsmm.graph_capture() as g:
        for c in chains:
            forward(c)
    assert _is_pair_kernel(xsmm.last_kernel(), f"mlp_chain_bf16_{n_chains}x{layers}layers_pair256x256"), xsmm.last_kernel()
    first = None
    for rep in range(3):
        for c in chains:
            for a in c[0][1:]:
                a.fill_(0x7FC0)
        n0 = xsmm.launch_count()
        g.launch()
        xsmm.sync()
        assert xsmm.launch_count() - n0 == 1, "all chains must share one launch"
        got = [[a.clone() for a in c[0][1:]] for c in chains]
        if first is None:
            first = got
            for gc, dc in zip(got, direct):
                for a, d in zip(gc, dc):
                    assert _ulp_diff(a, d) <= 1
        else:
            for gc, fc in zip(got, first):
                for a, f in zip(gc, fc):
                    assert torch.equal(a, f)
    for c in (0, n_chains // 2, n_chains - 1):
        Wc = [np.roll(w, c, axis=0) for w in hW]
        bc = [np.roll(b, c, axis=0) for b in hb]
        assert_close(BF16, first[c][-1].cpu().numpy().view(np.uint16), _oracle_mlp(xs[c], Wc, bc))
    g.destroy()


def test_large_batch_chain_is_cut_into_row_blocks_for_the_pairs():
    """One chain with m = 2048 batch rows per layer (BASELINE configs[4] on one GPU) plus one with m = 1024: rows are
    independent through the layers, so the runtime cuts the chains into 8 + 4 blocks of 256 rows, one CTA pair each."""
    import torch

    from tpp_mlir_b200 import xsmm

    layers = 3
    gen = oracle.TensorInit("normal", BF16, 77)

    def dev_t(a):
        return torch.from_numpy(a.view(np.int16)).cuda()

    specs = []
    for m in (2048, 1024):
        hW = [gen.fill(1024, 1024) for _ in range(layers)]
        hb = [gen.fill(1024) for _ in range(layers)]
        x = gen.fill(m, 1024)
        acts = [dev_t(x)] + [torch.zeros(m, 1024, dtype=torch.int16, device="cuda") for _ in range(layers)]
        h = xsmm.fused_brgemm_dispatch(BF16, m, 1024, 1024, 1024, 1024, 1024, 0, 0, 4, 0, 5, 4, 1)
        specs.append((m, h, x, hW, hb, acts, [dev_t(w) for w in hW], [dev_t(b) for b in hb]))
    with xsmm.graph_capture() as g:
        for m, h, x, hW, hb, acts, dW, db in specs:
            for l in range(layers):
                xsmm.fused_brgemm_invoke(BF16, h, acts[l], 0, dW[l], 0, acts[l + 1], 0, db[l], 0, 1)
    assert _is_pair_kernel(xsmm.last_kernel(), "mlp_chain_bf16_12x3layers_pair256x256"), xsmm.last_kernel()
    n0 = xsmm.launch_count()
    g.launch()
    xsmm.sync()
    assert xsmm.launch_count() - n0 == 1
    for m, h, x, hW, hb, acts, dW, db in specs:
        want = _oracle_mlp(x, hW, hb, m=m)
        assert_close(BF16, acts[-1].cpu().numpy().view(np.uint16), want)
    g.destroy()


def test_pair_kernel_without_bias_or_relu_and_with_padded_leading_dims():
    """Plain brgemm chains (no bias, no ReLU) and fused chains whose activations live in wider buffers (ld > n): the
    tensor maps carry lda / ldc, the epilogue must not touch the padding."""
    import torch

    from tpp_mlir_b200 import xsmm

    layers, n_chains, ld = 2, 12, 1024 + 64
    gen = oracle.TensorInit("normal", BF16, 5150)

    def dev_t(a):
        return torch.from_numpy(a.view(np.int16)).cuda()

    hW = [gen.fill(1024, 1024) for _ in range(layers)]
    dW = [dev_t(w) for w in hW]
    h = xsmm.brgemm_dispatch(BF16, 256, 1024, 1024, ld, 1024, ld, 0, 0, 4)
    xs, chains = [], []
    for c in range(n_chains):
        x = gen.fill(256, 1024)
        xs.append(x)
        acts = []
        for l in range(layers + 1):
            t = torch.full((256, ld), 0x1234, dtype=torch.int16, device="cuda")
            if l == 0:
                t[:, :1024] = dev_t(x)
            acts.append(t)
        chains.append(acts)
    with xsmm.graph_capture() as g:
        for acts in chains:
            for l in range(layers):
                xsmm.brgemm_invoke(BF16, h, acts[l], 0, dW[l], 0, acts[l + 1], 0, 1)
    assert _is_pair_kernel(xsmm.last_kernel(), f"mlp_chain_bf16_{n_chains}x{layers}layers_pair256x256"), xsmm.last_kernel()
    g.launch()
    xsmm.sync()
    for x, acts in zip(xs, chains):
        ref = x
        for W in hW:
            y = np.zeros((256, 1024), np.uint16)
            oracle.brgemm(BF16, 256, 1024, 1024, 1024, 1024, 1024, 0, 0, 4, ref, W, y, 1)
            ref = y
        out = acts[-1].cpu().numpy().view(np.uint16)
        assert_close(BF16, out[:, :1024].copy(), ref)
        for a in acts[1:]:
            assert (a[:, 1024:] == 0x1234).all(), "padding columns were overwritten"
    g.destroy()


# ---- tensor.pack / tensor.unpack: per-tile unary TPPs, batched into one kernel under graph capture ----
PACK_CASES = [
    # (dtype, M, N, bm, bn, outer_perm, unpack, transpose_tiles)
    (F32, 512, 1024, 32, 32, (0, 1), False, False),    # benchmarks/mlir/fp32-pack-gemm-operand-a-512x1024.mlir
    (F32, 1024, 512, 32, 32, (1, 0), False, False),    # fp32-pack-gemm-operand-b-512x1024.mlir
    (F32, 512, 512, 32, 32, (0, 1), True, False),      # fp32-unpack-gemm-operand-a-512x512.mlir
    (BF16, 256, 1024, 32, 32, (0, 1), False, False),   # the MLP's activation packing (mlir-gen --tiles=32,32,32)
    (BF16, 1024, 1024, 32, 32, (1, 0), False, False),  # ... and its weight packing
    (BF16, 256, 1024, 32, 32, (0, 1), True, False),
    (F32, 120, 200, 24, 40, (0, 1), False, False),     # rows that are no multiple of 16 bytes apart: scalar path
    (BF16, 96, 168, 12, 21, (1, 0), True, False),      # odd tile width
    (F32, 1024, 64, 512, 32, (0, 1), False, False),    # tall tiles: several TMA boxes along the rows of one tile
    (BF16, 64, 2048, 32, 1024, (0, 1), True, False),   # 2 KiB tile rows: wider than a TMA box, pointer-table kernel
    (BF16, 256, 512, 32, 64, (0, 1), False, True),     # tiles stored transposed (xsmm.unary transpose per tile)
    (F32, 128, 192, 32, 48, (1, 0), True, True),
]


@pytest.mark.parametrize("case", PACK_CASES, ids=lambda c: f"{'bf16' if c[0] == BF16 else 'f32'}-{c[1]}x{c[2]}-t{c[3]}x{c[4]}-"
                                                           f"{'perm' if c[5] == (1, 0) else 'id'}-"
                                                           f"{'unpack' if c[6] else 'pack'}{'-T' if c[7] else ''}")
def test_tiled_pack_unpack_is_one_batched_kernel_and_bit_exact(case):
    """The per-tile invoke sequence of a lowered tensor.pack / unpack, issued (a) directly - one launch per tile - and
    (b) inside a graph capture, where the runtime batches the run into ONE kernel. Both must equal the numpy restatement
    of the op bit for bit, and bytes outside the destination tiles must stay untouched."""
    import torch

    from tpp_mlir_b200 import harness, xsmm

    dtype, M, N, bm, bn, perm, unpack, tt = case
    npdt = np.float32 if dtype == F32 else np.uint16
    rng = np.random.default_rng(M * 7 + N)
    flat = (rng.standard_normal((M, N)).astype(np.float32) if dtype == F32
            else rng.integers(0, 65535, size=(M, N), dtype=np.uint16))
    packed = oracle.tensor_pack(flat, bm, bn, perm)
    if tt:
        packed = np.ascontiguousarray(packed.transpose(0, 1, 3, 2))
    tdt = torch.float32 if dtype == F32 else torch.int16

    def dev_t(a):
        return torch.from_numpy(a.view(np.int16) if dtype == BF16 else a).cuda()

    src_np, want = (packed, flat) if unpack else (flat, packed)
    src = dev_t(src_np)
    rp = harness.PackReplay(dtype, M, N, bm, bn, perm, unpack=unpack, transpose_tiles=tt)
    results = []
    for mode in ("direct", "captured"):
        dst = torch.zeros(want.size, dtype=tdt, device="cuda")
        args = (dst, src) if unpack else (src, dst)
        n0 = xsmm.launch_count()
        if mode == "direct":
            rp.run(*args)
            xsmm.sync()
            assert xsmm.launch_count() - n0 == rp.num_tiles
        else:
            with xsmm.graph_capture() as g:
                rp.run(*args)
            assert f"batch{rp.num_tiles}" in xsmm.last_kernel(), xsmm.last_kernel()
            n0 = xsmm.launch_count()
            g.launch()
            xsmm.sync()
            assert xsmm.launch_count() - n0 == 1, "the tiles of one pack must share one launch"
            dst.zero_()
            g.launch()   # replay
            xsmm.sync()
            g.destroy()
        got = dst.cpu().numpy().view(npdt).reshape(want.shape)
        assert np.array_equal(got, want), f"{mode}: pack/unpack result differs"
        results.append(got)


def test_full_size_pack_unpack_round_trip_is_one_tma_copy_each():
    """BASELINE-sized check of the batched tile moves: a 4096 x 4096 bf16 matrix packed into 32 x 32 tiles (16384 unary
    identity invokes) and unpacked again under capture. Each direction is ONE launch of the TMA-to-TMA grid copy
    (tile_grid.cu: the run walks a regular grid, no pointer table); packed == the numpy restatement, unpack(pack(x)) == x."""
    import torch

    from tpp_mlir_b200 import harness, xsmm

    M = N = 4096
    rng = np.random.default_rng(11)
    flat = rng.integers(0, 65535, size=(M, N), dtype=np.uint16)
    src = torch.from_numpy(flat.view(np.int16)).cuda()
    packed = torch.zeros(M * N, dtype=torch.int16, device="cuda")
    back = torch.zeros(M * N, dtype=torch.int16, device="cuda")
    fwd = harness.PackReplay(BF16, M, N, 32, 32, (0, 1))
    inv = harness.PackReplay(BF16, M, N, 32, 32, (0, 1), unpack=True)
    with xsmm.graph_capture() as g:
        fwd.run(src, packed)
        inv.run(back, packed)
    k_unpack = xsmm.last_kernel()
    assert "batch16384_tma128x128" in k_unpack, k_unpack
    n0 = xsmm.launch_count()
    g.launch()
    xsmm.sync()
    assert xsmm.launch_count() - n0 == 2, "one launch per direction"
    assert np.array_equal(packed.cpu().numpy().view(np.uint16).reshape(M // 32, N // 32, 32, 32),
                          oracle.tensor_pack(flat, 32, 32, (0, 1)))
    assert torch.equal(back.reshape(M, N), src)
    g.destroy()


def test_captured_tile_moves_respect_dependencies():
    """Tile copies that read what an earlier one wrote (a -> b -> c -> d -> e) or overwrite what an earlier one wrote
    must not be reordered into one batch; copies into disjoint tiles of ONE matrix (interleaved in address space) still
    batch. Mixed with BRGEMMs recorded during the same capture."""
    import torch

    from tpp_mlir_b200 import xsmm

    h = xsmm.unary_dispatch(xsmm.UNARY_IDENTITY, F32, 32, 32, 32, 32, 0)
    bufs = [torch.zeros(32 * 32, device="cuda") for _ in range(6)]
    bufs[0].copy_(torch.arange(1024, dtype=torch.float32))
    other = torch.full((1024,), 5.0, device="cuda")
    with xsmm.graph_capture() as g:
        for i in range(5):
            xsmm.unary_invoke(F32, h, bufs[i], 0, bufs[i + 1], 0)       # chain: each reads the previous output
        xsmm.unary_invoke(F32, h, other, 0, bufs[1], 0)                  # WAW + WAR on bufs[1]
    n0 = xsmm.launch_count()
    g.launch()
    xsmm.sync()
    assert xsmm.launch_count() - n0 == 6, "dependent copies must stay separate launches"
    assert torch.equal(bufs[5], bufs[0]) and torch.equal(bufs[1], other)
    g.destroy()
    # 16 tiles of one 128 x 128 matrix copied into another one: bounding ranges interleave, elements are disjoint
    ht = xsmm.unary_dispatch(xsmm.UNARY_IDENTITY, F32, 32, 32, 128, 128, 0)
    a = torch.arange(128 * 128, dtype=torch.float32, device="cuda")
    b = torch.zeros(128 * 128, device="cuda")
    with xsmm.graph_capture() as g:
        for i in range(4):
            for j in range(4):
                off = i * 32 * 128 + j * 32
                xsmm.unary_invoke(F32, ht, a, off, b, off)
    assert "batch16" in xsmm.last_kernel(), xsmm.last_kernel()
    g.launch()
    xsmm.sync()
    assert torch.equal(a, b)
    g.destroy()


def test_pair_kernel_chain_with_layers_of_different_width():
    """A funnel MLP 1024 -> 512 -> 256 -> 512 (three different dispatches per chain): the number of 256-column output
    tiles changes from layer to layer, and so does the number of per-tile hand-off barriers each layer boundary uses."""
    import torch

    from tpp_mlir_b200 import xsmm

    sizes = (1024, 512, 256, 512)
    n_chains = 14
    gen = oracle.TensorInit("normal", BF16, 4242)

    def dev_t(a):
        return torch.from_numpy(a.view(np.int16)).cuda()

    hs = [xsmm.fused_brgemm_dispatch(BF16, 256, n, k, k, n, n, 0, 0, 4, 0, 5, 4, 1) for k, n in zip(sizes[:-1], sizes[1:])]
    chains = []
    for _ in range(n_chains):
        Ws = [gen.fill(k, n) for k, n in zip(sizes[:-1], sizes[1:])]
        bs = [gen.fill(n) for n in sizes[1:]]
        x = gen.fill(256, sizes[0])
        acts = [dev_t(x)] + [torch.zeros(256, n, dtype=torch.int16, device="cuda") for n in sizes[1:]]
        chains.append((x, Ws, bs, acts, [dev_t(w) for w in Ws], [dev_t(b) for b in bs]))
    with xsmm.graph_capture() as g:
        for x, Ws, bs, acts, dW, db in chains:
            for l, h in enumerate(hs):
                xsmm.fused_brgemm_invoke(BF16, h, acts[l], 0, dW[l], 0, acts[l + 1], 0, db[l], 0, 1)
    assert _is_pair_kernel(xsmm.last_kernel(), f"mlp_chain_bf16_{n_chains}x3layers_pair256x256"), xsmm.last_kernel()
    for rep in range(2):
        for c in chains:
            for a in c[3][1:]:
                a.fill_(0x7FC0)
        g.launch()
        xsmm.sync()
        for x, Ws, bs, acts, dW, db in chains:
            ref = x
            for W, b in zip(Ws, bs):
                y = np.zeros((256, W.shape[1]), np.uint16)
                oracle.fused_brgemm(BF16, 256, W.shape[1], W.shape[0], W.shape[0], W.shape[1], W.shape[1], 0, 0, 4, 0, 5,
                                    4, 1, ref, W, y, b, 1)
                ref = y
            assert_close(BF16, acts[-1].cpu().numpy().view(np.uint16), ref)
    g.destroy()
