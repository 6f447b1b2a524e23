"""CPU-only: tpp_run_standin (tpp_mlir_b200/csrc/harness/tpp_run_standin.cpp - the native stand-in for
`mlir-gen ... | tpp-run -n N`) against a recording stub of the C-ABI (tests/stubs/xsmm_abi_stub_standin.cpp). Checks WHAT
it asks the runtime to do for the reference's benchmark command lines: dispatch arguments, the Appendix-B invoke stream,
which buffers are registered / marked temporary, graph use, and mlir-gen's FLOP count."""
import json
import os
import subprocess

import pytest

from tpp_mlir_b200 import harness

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "reference_bench_shapes.json")) as _f:
    SHAPES = json.load(_f)


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("standin_stub") / "tpp_run_standin"   # the name bench_driver --build looks for
    cmd = ["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tpp_mlir_b200", "csrc", "harness", "tpp_run_standin.cpp"),
           os.path.join(ROOT, "tests", "stubs", "xsmm_abi_stub_standin.cpp"), "-o", str(out)]
    subprocess.run(cmd, check=True)
    return str(out)


def run(exe, tmp_path, *args):
    log = tmp_path / "stub.json"
    env = dict(os.environ, STUB_LOG=str(log))
    env.pop("TPP_XSMM_VNNI", None)
    r = subprocess.run([exe, *args], capture_output=True, text=True, timeout=120, env=env)
    if r.returncode != 0:
        return r.returncode, None, None
    return 0, json.loads(r.stdout.strip().splitlines()[-1]), json.loads(log.read_text())


def mlir_gen_line(s):
    parts = [f"--kernel={s['kernel']}"]
    if s["bias"]:
        parts.append("--bias")
    if s["relu"]:
        parts.append("--relu")
    parts.append(f"--float-type={s['float_type']}")
    if s["vnni"]:
        parts.append(f"--vnni={s['vnni']}")
    parts += [f"--batch={s['batch']}", "--layers=" + ",".join(map(str, s["layers"]))]
    if s["tiles"]:
        parts.append("--tiles=" + ",".join(map(str, s["tiles"])))
    return " ".join(parts)


def test_default_is_the_headline_mlp_with_the_reference_tiling(exe, tmp_path):
    """no options: mlir-gen --kernel=const --bias --relu --float-type=bf16 --batch=256 --layers=1024,1024,1024,1024
    --tiles=32,32,32 --vnni=2 (benchmarks/config/omp/mlir-bf16.json:37) with plain host pointers"""
    rc, out, log = run(exe, tmp_path, "-n", "3")
    assert rc == 0
    gflags = 4 | 2048 | 64 | 128   # beta_0, vnni_b, the two AMX tile-config bits
    assert log["dispatches"] == [[1, 2, 32, 32, 32, 32, 32, 32, 1024, 1024, gflags, 0, 5, 4, 1],
                                 [2, 2, 32, 32, 32, 32, 32, 32, 1024, 1024, gflags]]
    assert log["fused_invokes"] == (1 + 3) * 768 and log["tilecfg_invokes"] == 2 * log["fused_invokes"]
    assert log["registered"] == 0 and log["temporaries"] == 0 and log["graphs"] == 0 and log["timers"] == 1
    # (iN, iK) = (0, 0), (0, 1), ...: A stays, B steps by a column block of 32 x 1024 weights,
    # C by one 32 x 32 tile, the bias by 32; batch-reduce count = 1024 / 32
    assert log["first_invokes"][:3] == [[1, 2, 0, 0, 0, 0, 1, 32], [1, 2, 0, 32768, 1024, 32, 1, 32],
                                        [1, 2, 0, 65536, 2048, 64, 1, 32]]
    assert out["total_flops"] == 1612185600 and out["mode"] == "strict"


def test_modes_register_mark_and_capture(exe, tmp_path):
    base = ["--batch", "256", "--layers", "1024,1024,1024,1024", "--tiles", "256,1024,1024", "--vnni", "0", "-n", "100"]
    rc, _, dev = run(exe, tmp_path, *base, "--mode", "device")
    assert rc == 0
    # 3 weights + 3 biases + 4 activation buffers; the two intermediate activations are function-local temporaries
    assert dev["registered"] == 10 and dev["temporaries"] == 2 and dev["temporary_bytes"] == 2 * 256 * 1024 * 2
    assert dev["graphs"] == 0 and dev["fused_invokes"] == (1 + 100) * 3 and dev["lazy"] == 0
    rc, _, lazy = run(exe, tmp_path, *base, "--mode", "lazy")
    assert rc == 0 and lazy["lazy"] == 1 and lazy["graphs"] == 0
    rc, _, gr = run(exe, tmp_path, *base, "--mode", "graph")
    assert rc == 0
    # the body is recorded once; warm-up and timed iterations are launches of that graph (patches/0005)
    assert gr["graphs"] == 1 and gr["captured_invokes"] == 3 and gr["fused_invokes"] == 3 and gr["graph_launches"] == 1 + 100
    rc, _, args_mode = run(exe, tmp_path, *base, "--mode", "device", "--kernel", "args")
    assert rc == 0 and args_mode["temporaries"] == 0   # --kernel=args: every layer's output is a kernel argument


@pytest.mark.parametrize("entry", SHAPES, ids=lambda e: mlir_gen_line(e["shape"]).replace(" ", ""))
def test_reference_benchmark_lines_become_the_right_call_streams(entry, exe, tmp_path):
    """every distinct mlir-gen line of benchmarks/config/{fc,matmul,omp,base}/*.json"""
    s = entry["shape"]
    rc, out, log = run(exe, tmp_path, "--mlir-gen", mlir_gen_line(s), "-n", "2", "--mode", "graph")
    assert rc == 0
    bf16 = s["float_type"] == "bf16"
    dtype = 2 if bf16 else 1
    vnni = (s["vnni"] or 2) if bf16 else 0
    bn, bk, bc = s["tiles"] or (32, 32, 32)
    fused = s["bias"] or s["relu"]
    gflags = 4 | (2048 if vnni else 0) | ((64 | 128) if bf16 else 0)
    geom = [dtype, bn, bk, bc, bc, bk, bk, bn * bc, bc * bk, gflags]
    want = [[1, *geom, 0, 5 if s["relu"] else 0, 4 if s["bias"] else 0, 1 if s["bias"] else 0]] if fused else [[0, *geom]]
    if bf16:
        want.append([2, *geom])
    assert log["dispatches"] == want
    layers, batch = s["layers"], s["batch"]
    invokes = sum((batch // bn) * (k // bk) for k in layers[1:])
    assert log["fused_invokes" if fused else "brgemm_invokes"] == invokes == log["captured_invokes"]
    assert log["brgemm_invokes" if fused else "fused_invokes"] == 0
    assert log["tilecfg_invokes"] == (2 * invokes if bf16 else 0)
    assert log["graphs"] == 1 and log["graph_launches"] == 1 + 2
    first = log["first_invokes"][0]
    assert first == [1 if fused else 0, dtype, 0, 0, 0, 0, 1 if s["bias"] else 0, layers[0] // bc]
    # the last tile of every operand ends exactly at the end of its buffer (largest layer bounds the maxima)
    off_a, off_b, off_c, off_d = log["max_off"]
    assert off_a == max((batch // bn - 1) * (c // bc) * bn * bc for c in layers[:-1])
    assert off_b == max((k // bk - 1) * (c // bc) * bc * bk for c, k in zip(layers[:-1], layers[1:]))
    assert off_c == max(batch * k - bn * bk for k in layers[1:])
    assert off_d == (max(k - bk for k in layers[1:]) if s["bias"] else 0)
    n_bufs = 2 * (len(layers) - 1) + len(layers)
    assert log["registered"] == n_bufs
    assert log["temporaries"] == (len(layers) - 2 if s["kernel"] == "const" else 0)
    assert log["vnni_env_at_dispatch"] == ("4" if vnni == 4 else "")
    cfg = harness.MlpConfig(batch=batch, layers=tuple(layers), tiles=(bn, bk, bc), dtype=dtype, vnni=bool(vnni),
                            bias=s["bias"], relu=s["relu"])
    assert out["total_flops"] == cfg.flops()
    assert out["float_type"] == s["float_type"] and out["vnni"] == vnni


def test_flops_follow_the_reference_count(exe, tmp_path):
    # test/Integration/mlir-gen-fc.mlir:1-5 and mlir-gen-matmul.mlir:1-7: BENCH_TOTAL_FLOPS 453181440 / 452984832
    line = "--kernel=args {} --seed=0 --float-type=f32 --batch=128 --layers=2304,768 --tiles=64,48,64"
    rc, out, _ = run(exe, tmp_path, "--mlir-gen", line.format("--bias --relu"), "-n", "1")
    assert rc == 0 and out["total_flops"] == 453181440
    rc, out, _ = run(exe, tmp_path, "--mlir-gen", line.format(""), "-n", "1")
    assert rc == 0 and out["total_flops"] == 452984832


@pytest.mark.parametrize("args", [("--mlir-gen", "--kernel=args --float-type=f16 --batch=128 --layers=64,64"),
                                  ("--mlir-gen", "--batch=100 --layers=64,64"),           # 100 rows: no multiple of 32
                                  ("--mlir-gen", "--batch=128 --layers=64,64 --softmax"),   # not part of this path
                                  ("--float-type", "f32", "--vnni", "2"),
                                  ("--layers", "1024,1024,1024", "--tiles", "32,64,32"),
                                  ("--vnni", "3")])
def test_bad_command_lines_are_refused(args, exe, tmp_path):
    rc, _, _ = run(exe, tmp_path, *args)
    assert rc == 2


# ---- tpp_mlir_b200/bench_driver.py: the reference's benchmarks/driver.py contract on top of the stand-in -------------------
DRIVER_CONFIG = [
    {"mlp_bf16_dp2_mlir": {
        "bf16_dp2_3x1024_omp_16_mlir": {   # benchmarks/config/omp/mlir-bf16.json:34-62
            "type": "IR-GEN",
            "benchmark": ["mlir-gen", "--kernel=const --bias --relu --float-type=bf16 --vnni=2 --batch=256 "
                                      "--layers=1024,1024,1024,1024 --tiles=32,32,32"],
            "environment": {"OMP_NUM_THREADS": "16"},
            "flags": ["-n", "100", "-run-args='-def-parallel'"],
            "extensions": [],
        },
        "never_on_this_host": {
            "type": "IR-GEN",
            "benchmark": ["mlir-gen", "--kernel=const --float-type=bf16 --batch=256 --layers=1024,1024"],
            "environment": {}, "flags": ["-n", "100"], "extensions": ["(no_such_cpu_flag)"],
        }}},
    {"fc_1024x352x512": {
        "fc_fp32_single_dnn": {   # benchmarks/config/fc/1024x352x512.json
            "type": "XSMM-DNN", "benchmark": "xsmm_dnn_mlp", "environment": {"OMP_NUM_THREADS": "1"},
            "flags": ["100", "1024", "3", "F", "32", "32", "32", "0", "1", "512", "352"], "extensions": [],
        },
        "fc_fp32_single_mlir": {
            "type": "IR-GEN",
            "benchmark": ["mlir-gen", "--kernel=args --bias --relu --float-type=f32 --batch=1024 --layers=512,352 --tiles=32,32,32"],
            "environment": {"OMP_NUM_THREADS": "1"}, "flags": ["-n", "100"], "extensions": [],
        },
        "from_a_file": {
            "type": "MLIR", "benchmark": "fp32-mha-tensorflow.mlir", "environment": {}, "flags": ["-n", "100"], "extensions": [],
        }}},
]


def _drive(exe, tmp_path, config, *extra):
    import sys

    path = tmp_path / "config.json"
    path.write_text(json.dumps(config))
    env = dict(os.environ, PYTHONPATH=ROOT)
    return subprocess.run([sys.executable, "-m", "tpp_mlir_b200.bench_driver", "-c", str(path), "--build", os.path.dirname(exe),
                           *extra], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)


def test_bench_driver_prints_the_reference_report(exe, tmp_path):
    """`Benchmark: <name>` / `<run:28>: <%9.3f> gflops` / blank line (benchmarks/driver.py:498-515, controller.py:316);
    entries whose extensions the host lacks are dropped, run types that need the reference's toolchain are skipped; the
    stub's timer reports 1 s for the timed loop, so gflops = FLOPs * n / 1e9."""
    r = _drive(exe, tmp_path, DRIVER_CONFIG, "-n", "10")
    assert r.returncode == 0, r.stderr
    assert r.stdout == ("Benchmark: mlp_bf16_dp2_mlir\n"
                        f"{'bf16_dp2_3x1024_omp_16_mlir':28}: {1612185600 * 10 / 1e9:9.3f} gflops\n"
                        "\n"
                        "Benchmark: fc_1024x352x512\n"
                        f"{'fc_fp32_single_mlir':28}: {(2 * 1024 * 512 * 352 + 2 * 1024 * 352) * 10 / 1e9:9.3f} gflops\n"
                        "\n")


def test_bench_driver_iterations_modes_and_json(exe, tmp_path):
    r = _drive(exe, tmp_path, DRIVER_CONFIG[:1], "--json", "--mode", "device", "--ignore-extensions")
    assert r.returncode == 0, r.stderr
    rows = [json.loads(line) for line in r.stdout.splitlines() if line.startswith("{")]
    assert [row["run"] for row in rows] == ["bf16_dp2_3x1024_omp_16_mlir", "never_on_this_host"]
    assert all(row["mode"] == "device" and row["iterations"] == 100 for row in rows)   # -n from the entry's own flags
    assert rows[1]["float_type"] == "bf16" and rows[1]["vnni"] == 2 and rows[1]["bias"] == 0   # mlir-gen's defaults


def test_bench_driver_errors(exe, tmp_path):
    bad = [{"broken": {"run": {"type": "IR-GEN", "benchmark": ["mlir-gen", "--batch=100 --layers=64,64"], "environment": {},
                               "flags": [], "extensions": []}}}]
    assert _drive(exe, tmp_path, bad).returncode == 1
    r = _drive(exe, tmp_path, bad, "--ignore-errors")
    assert r.returncode == 0 and r.stdout == "Benchmark: broken\n\n"
    assert _drive(exe, tmp_path, [{"a": {}, "b": {}}]).returncode == 1   # one benchmark per list element
