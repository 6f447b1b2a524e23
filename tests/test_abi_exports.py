"""CPU-only: the drop-in library builds for sm_100a, loads, and exports every symbol that
include/tpp_xsmm_abi.h declares (and the 13+2+1 names tpp-mlir binds). No compute calls."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REFERENCE_SYMBOLS = [  # runtime/Xsmm/XsmmRunnerUtils.h:22-83, runtime/PerfRunnerUtils.h:22-24, VNNIUtils.cpp:36
    "xsmm_gemm_dispatch", "xsmm_unary_dispatch", "xsmm_binary_dispatch", "xsmm_brgemm_dispatch",
    "xsmm_fused_brgemm_dispatch", "xsmm_intel_amx_tile_config_dispatch", "xsmm_gemm_invoke", "xsmm_unary_invoke",
    "xsmm_unary_scalar_invoke", "xsmm_binary_invoke", "xsmm_brgemm_invoke", "xsmm_fused_brgemm_invoke",
    "xsmm_intel_amx_tile_config_invoke", "perf_start_timer", "perf_stop_timer", "libxsmm_cpuid_dot_pack_factor",
]


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "tpp_xsmm_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"TPP_XSMM_EXPORT\s+[\w\s\*]+?\b(\w+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from tpp_mlir_b200 import _build

    return ctypes.CDLL(_build.build())


def test_header_declares_the_reference_abi():
    decl = declared_symbols()
    for s in REFERENCE_SYMBOLS:
        assert s in decl, s


def test_library_exports_every_declared_symbol(lib):
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/tpp_xsmm_abi.h but not exported"


def test_python_binding_table_matches_header():
    from tpp_mlir_b200 import xsmm

    assert sorted(xsmm.EXPORTS) == declared_symbols()


def test_no_cpu_symbols_or_oracle_linked(lib):
    # the product must not carry the oracle (xo_*) or any host compute fallback
    from tpp_mlir_b200 import _build

    out = subprocess.run(["nm", "-D", "--defined-only", _build.lib_path()], capture_output=True, text=True).stdout
    assert "xo_" not in out and "ti_fill" not in out


def test_sass_has_blackwell_tensor_and_tma_instructions():
    from tpp_mlir_b200 import _build

    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not installed")
    sass = subprocess.run([cuobjdump, "-sass", _build.lib_path()], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass      # tcgen05.mma
    assert "LDTM" in sass         # tcgen05.ld
    assert "UTMALDG" in sass      # cp.async.bulk.tensor
    assert "HMMA.16816" not in sass  # no legacy mma.sync path


def test_pack_factor_and_timers_work_without_gpu(lib):
    lib.libxsmm_cpuid_dot_pack_factor.restype = ctypes.c_int
    assert lib.libxsmm_cpuid_dot_pack_factor(2) == 2   # bf16 -> VNNI-2
    assert lib.libxsmm_cpuid_dot_pack_factor(1) == 1
    lib.perf_start_timer.restype = ctypes.c_int64
    lib.perf_stop_timer.restype = ctypes.c_double
    lib.perf_stop_timer.argtypes = [ctypes.c_int64]
    t0 = lib.perf_start_timer()
    assert 0.0 <= lib.perf_stop_timer(t0) < 5.0


def test_dispatch_without_gpu_fails_loudly():
    """No CPU fallback: on a box without a GPU the first dispatch must exit non-zero with a diagnostic."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    code = ("from tpp_mlir_b200 import xsmm; "
            "xsmm.brgemm_dispatch(2, 32, 32, 32, 32, 32, 32, 1024, 1024, 4); print('SURVIVED')")
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert p.returncode != 0
    assert "SURVIVED" not in p.stdout
    assert "no CUDA device" in p.stderr


def _param_class(p):
    p = p.strip()
    if "*" in p:
        return "ptr"
    if re.match(r"(const\s+)?float\b", p):
        return "float"
    return "i64"   # int64_t and the libxsmm enums the lowering passes as i64 (ConvertXsmmToFunc.cpp:37-78)


def _param_name(p):
    m = re.search(r"(\w+)$", p.strip())
    return m.group(1) if m and not re.match(r"(int64_t|float|libxsmm_\w+)$", m.group(1)) else None


def test_prototypes_follow_the_reference_header_and_the_pinned_call_order():
    """Every one of the 13 xsmm_* entry points has the reference's result type, parameter count and parameter classes
    (i64 / pointer / float, in order; runtime/Xsmm/XsmmRunnerUtils.h:22-83, extracted into
    tests/golden/reference_abi_calls.json) and, where the reference names a parameter, the same name; the argument counts
    the lowering's FileCheck lines pin (test/Conversion/XsmmToFunc/xsmm-to-func.mlir) agree too."""
    import json

    with open(os.path.join(ROOT, "tests", "golden", "reference_abi_calls.json")) as f:
        ref = json.load(f)
    text = open(os.path.join(ROOT, "include", "tpp_xsmm_abi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    ours = {}
    for m in re.finditer(r"TPP_XSMM_EXPORT\s+(\w+)\s+(\w+)\s*\(([^;]*?)\)\s*;", text, re.S):
        ours[m.group(2)] = (m.group(1), [re.sub(r"\s+", " ", p.strip()) for p in m.group(3).split(",") if p.strip()])
    assert len(ref["header"]) == 13
    for name, proto in ref["header"].items():
        assert name in ours, name
        result, params = ours[name]
        assert result == proto["result"], name
        assert [_param_class(p) for p in params] == [_param_class(p) for p in proto["params"]], name
        for mine, theirs in zip(params, proto["params"]):
            want = _param_name(theirs)
            if want and want not in ("dType", "data_type", "unary_op_type", "binary_op_type"):
                assert _param_name(mine) == want, (name, mine, theirs)
    for name, lines in ref["calls"].items():
        for c in lines:
            assert c["num_args"] == len(ours[name][1]), (name, c["line"])


def test_nvtx_ranges_need_no_profiler_and_no_gpu():
    """TPP_XSMM_NVTX=1 wraps every invoke entry in an NVTX range (runtime.cu: NvtxRange). Without a profiler attached the
    NVTX calls are no-ops: an invoke with a bogus handle must end the way it does without the variable - the reference's
    error convention (message on stderr, exit(-1): XsmmRunnerUtils.cpp:132-137) - not in a crash."""
    code = "\n".join([
        "import ctypes, sys",
        "sys.path.insert(0, %r)" % ROOT,
        "from tpp_mlir_b200 import _build",
        "lib = ctypes.CDLL(_build.lib_path())",
        "lib.xsmm_unary_invoke.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64]",
        "lib.xsmm_unary_invoke(1, 0, None, 0, None, 0)",
    ])
    outs = []
    for nvtx in ("0", "1"):
        env = dict(os.environ, TPP_XSMM_NVTX=nvtx)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, env=env)
        outs.append((r.returncode, r.stderr.strip().splitlines()[-1]))
    assert outs[0] == outs[1]
    assert outs[0][0] == 255 and "handle" in outs[0][1]
