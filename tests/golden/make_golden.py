#!/usr/bin/env python
"""Extract the known-answer vectors of the reference's own tests into a fixture.

Run in the build container (needs /root/reference, which does not exist on the
GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/reference_vectors.json. Every number in that file is copied
out of a reference test file (dense<> constants and FileCheck CHECK/RESULT lines);
nothing is computed here. The dispatch arguments that go with each vector are
written in tests/golden_cases.py next to the reference file:line they come from.
"""
from __future__ import annotations

import json
import os
import re
import sys

REF = os.environ.get("TPP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")

NUM = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?"


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def nums(s):
    return [float(x) for x in re.findall(NUM, s)]


def dense_blocks(text):
    """All `dense<[ ... ]>` literals of a file, in order, as flat float lists."""
    out = []
    i = 0
    while True:
        i = text.find("dense<[", i)
        if i < 0:
            break
        depth = 0
        j = i + len("dense<")
        start = j
        while True:
            c = text[j]
            if c == "[":
                depth += 1
            elif c == "]":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        out.append(nums(text[start:j + 1]))
        i = j
    return out


def check_lines(text, prefix="CHECK"):
    """Numbers of every `// PREFIX[-SAME|-COUNT-n]: ...` line, one list per line, with the
    repeat count of CHECK-COUNT-n."""
    rows = []
    for line in text.splitlines():
        m = re.match(r"\s*//\s*" + prefix + r"(-SAME|-COUNT-(\d+))?:\s*(.*)$", line)
        if not m:
            continue
        body = m.group(3)
        if "(" not in body:
            continue
        vals = nums(body)
        rep = int(m.group(2)) if m.group(2) else 1
        rows.append({"values": vals, "repeat": rep})
    return rows


def flat_checks(text, prefix="CHECK"):
    out = []
    for r in check_lines(text, prefix):
        out.extend(r["values"] * r["repeat"])
    return out


def main():
    g = {}

    def add(name, src, **kw):
        kw["source"] = src
        g[name] = kw

    t = read("test/Integration/xsmm-brgemm.mlir")
    add("brgemm_f32_ones", "test/Integration/xsmm-brgemm.mlir:8-20", expected=flat_checks(t))

    t = read("test/BF16/Integration/xsmm-brgemm-bf16.mlir")
    add("brgemm_bf16_vnni", "test/BF16/Integration/xsmm-brgemm-bf16.mlir:5-20", expected=flat_checks(t))

    t = read("test/BF16/Integration/xsmm-ternary-bf16.mlir")
    add("brgemm_bf16_vnni_batch64", "test/BF16/Integration/xsmm-ternary-bf16.mlir:5-19", expected=flat_checks(t))

    t = read("test/BF16/Integration/xsmm-gemm-bf16.mlir")
    add("gemm_bf16_vnni", "test/BF16/Integration/xsmm-gemm-bf16.mlir:5-17", expected=flat_checks(t))

    for name, rel in (("fused_f32", "test/Integration/xsmm-quarternary.mlir"),
                      ("fused_bf16_vnni", "test/BF16/Integration/xsmm-quarternary-bf16.mlir")):
        t = read(rel)
        m = re.search(r"%outVal = arith.constant (" + NUM + ")", t)
        thr = re.search(r"%threshold = arith.constant (" + NUM + ")", t)
        add(name, rel + ":4-16", expected_fill=float(m.group(1)), threshold=float(thr.group(1)))

    t = read("test/Integration/xsmm-fusion.mlir")
    add("fused_f32_seed123", "test/Integration/xsmm-fusion.mlir:5-57", expected=flat_checks(t, "RESULT"))

    t = read("test/Integration/xsmm-transpose.mlir")
    add("transpose_f32", "test/Integration/xsmm-transpose.mlir:5-41", input=dense_blocks(t)[0], expected=flat_checks(t))

    t = read("test/Integration/transpose-bf16.mlir")
    add("vnni2_bf16_seed123", "test/Integration/transpose-bf16.mlir:1-35", expected=flat_checks(t))

    t = read("test/BF16/Integration/vnni-packing.mlir")
    add("vnni2_pack_16x16", "test/BF16/Integration/vnni-packing.mlir:5-41", input=dense_blocks(t)[0],
        expected_prefix=flat_checks(t))

    t = read("test/BF16/Integration/vnni-packing-chain.mlir")
    d = dense_blocks(t)
    add("vnni2_pack_chain", "test/BF16/Integration/vnni-packing-chain.mlir:7-66", input=d[0], expected=d[1])

    t = read("test/Integration/xsmm-unary.mlir")
    add("unary_relu_f32", "test/Integration/xsmm-unary.mlir:5-13", expected=flat_checks(t))
    t = read("test/BF16/Integration/xsmm-unary-bf16.mlir")
    add("unary_relu_bf16", "test/BF16/Integration/xsmm-unary-bf16.mlir", expected=flat_checks(t))
    t = read("test/Integration/xsmm-zero.mlir")
    add("unary_zero_f32", "test/Integration/xsmm-zero.mlir:5-15", expected=flat_checks(t))
    t = read("test/BF16/Integration/xsmm-zero-bf16.mlir")
    add("unary_zero_bf16", "test/BF16/Integration/xsmm-zero-bf16.mlir", expected=flat_checks(t))
    t = read("test/Integration/xsmm-binary.mlir")
    add("binary_add_f32", "test/Integration/xsmm-binary.mlir:5-17", expected=flat_checks(t))
    t = read("test/BF16/Integration/xsmm-binary-bf16.mlir")
    add("binary_add_bf16", "test/BF16/Integration/xsmm-binary-bf16.mlir", expected=flat_checks(t))

    for op in ("mul", "sub"):
        t = read(f"test/Integration/xsmm-{op}.mlir")
        add(f"binary_{op}_f32", f"test/Integration/xsmm-{op}.mlir", input=dense_blocks(t)[0], expected=flat_checks(t))
    t = read("test/Integration/xsmm-div.mlir")
    d = dense_blocks(t)
    add("binary_div_f32", "test/Integration/xsmm-div.mlir", lhs=d[0], rhs_full=d[1], rhs_col=d[2], rhs_row=d[3],
        rhs_scalar=d[4], expected_all=flat_checks(t))

    t = read("test/Integration/xsmm-strided-brgemm.mlir")
    d = dense_blocks(t)
    add("strided_brgemm", "test/Integration/xsmm-strided-brgemm.mlir:17-109", D=d[0], A=d[1], B=d[2],
        expected=flat_checks(t))
    t = read("test/Integration/xsmm-strided-brgemm1.mlir")
    d = dense_blocks(t)
    add("strided_gemm1", "test/Integration/xsmm-strided-brgemm1.mlir:15-100", A=d[0], B=d[1], expected=flat_checks(t))

    for i in (2, 3):
        t = read(f"test/Integration/xsmm-strided-brgemm{i}.mlir")
        d = dense_blocks(t)
        add(f"strided_gemm{i}", f"test/Integration/xsmm-strided-brgemm{i}.mlir", A=d[0], B=d[1], expected=flat_checks(t))

    t = read("test/BF16/Integration/mlir-gen-bf16.mlir")
    mm = re.search(r"GEN-MATMUL-BF16: \(([^)]*)\)", t)
    fc = re.search(r"GEN-FC-BF16: \(([^)]*)\)", t)
    add("mlir_gen_bf16", "test/BF16/Integration/mlir-gen-bf16.mlir:9-26", matmul_row=nums(mm.group(1)), fc_row=nums(fc.group(1)),
        batch=16, layers=[16, 16])

    t = read("test/BF16/Integration/mlp-all-bf16-tpprun.mlir")
    m = re.search(r"%c4 = arith.constant (" + NUM + ")", t)
    thr = re.search(r"%threshold = arith.constant (" + NUM + ")", t)
    add("mlp_all_ones_bf16", "test/BF16/Integration/mlp-all-bf16-tpprun.mlir:4-135", expected_fill=float(m.group(1)),
        threshold=float(thr.group(1)), batch=128, layers=[256, 512, 1024, 2048, 1000])

    t = read("test/Integration/matmul_64x64x64.mlir")
    add("matmul_64x64x64_f32", "test/Integration/matmul_64x64x64.mlir:1-16", raw_checks=check_lines(t))

    # round 2, second batch: more of the reference's tests that end in xsmm invokes of this path
    t = read("test/Integration/xsmm-ternary.mlir")
    add("brgemm_f32_ternary", "test/Integration/xsmm-ternary.mlir:5-16", expected=flat_checks(t))
    for name in ("matmul_48x64x96", "matmul_64x48x96"):
        t = read(f"test/Integration/{name}.mlir")
        add(f"{name}_f32", f"test/Integration/{name}.mlir", raw_checks=check_lines(t))
    t = read("test/Integration/tpp-brgemm-non-unit-batch.mlir")
    d = dense_blocks(t)
    add("brgemm_f32_non_unit_batch", "test/Integration/tpp-brgemm-non-unit-batch.mlir:12-66", A=d[0], B=d[1],
        expected=flat_checks(t))
    t = read("test/Integration/transpose-fp32.mlir")
    add("transpose_f32_seed123", "test/Integration/transpose-fp32.mlir:1-25", expected=flat_checks(t))
    t = read("test/BF16/Integration/matmul-pbf16.mlir")
    add("matmul_pbf16", "test/BF16/Integration/matmul-pbf16.mlir:9-41", expected=flat_checks(t))
    t = read("test/BF16/Integration/mlp-single-layer-blocked-bf16.mlir")
    add("mlp_single_layer_blocked_bf16", "test/BF16/Integration/mlp-single-layer-blocked-bf16.mlir:11-57",
        expected=flat_checks(t))
    t = read("test/Integration/tpp-matmul.mlir")
    d = dense_blocks(t)
    add("tpp_matmul_f32", "test/Integration/tpp-matmul.mlir:13-58", A=d[0], B=d[1], expected=flat_checks(t))
    t = read("test/Integration/tpp-relu.mlir")
    add("tpp_relu_f32", "test/Integration/tpp-relu.mlir:14-54", input=dense_blocks(t)[0], expected=flat_checks(t))
    t = read("test/Integration/copy.mlir")
    d = dense_blocks(t)
    add("copy_broadcasts_f32", "test/Integration/copy.mlir:18-146", row=d[0], col=d[1], scalar=d[-1], expected_all=flat_checks(t))
    t = read("test/Integration/mlp-fp32-1layer-512.mlir")
    add("mlp_fp32_1layer_512", "test/Integration/mlp-fp32-1layer-512.mlir:8-31", raw_checks=check_lines(t))

    # round 2, third batch: tests whose operands are dense literals / fills and whose printed tensors are the
    # answers of gemm, bias + relu, binary add, relu and tile-wise pack / unpack sequences
    t = read("test/Integration/tpp-brgemm.mlir")
    d = dense_blocks(t)
    add("tpp_brgemm_f32", "test/Integration/tpp-brgemm.mlir:12-57", A=d[0], B=d[1], expected=flat_checks(t))
    t = read("test/Integration/simple-gemm.mlir")
    add("simple_gemm_f32", "test/Integration/simple-gemm.mlir:5-11", expected=flat_checks(t))
    t = read("test/Integration/packed-matmul.mlir")
    d = dense_blocks(t)
    add("packed_matmul_f32", "test/Integration/packed-matmul.mlir:25-87", A=d[0], B=d[1], bias=d[2],
        expected=flat_checks(t))
    t = read("test/Integration/tpp-run-xsmm-path.mlir")
    add("xsmm_path_add_f32", "test/Integration/tpp-run-xsmm-path.mlir:8-38", expected=flat_checks(t))
    t = read("test/Integration/matmul-tpp-with-print.mlir")
    add("matmul_tpp_with_print_f32", "test/Integration/matmul-tpp-with-print.mlir:17-59", expected=flat_checks(t))
    t = read("test/Integration/result-out-arg.mlir")
    d = dense_blocks(t)
    add("result_out_arg_f32", "test/Integration/result-out-arg.mlir:9-48", A=d[0], B=d[1], expected=flat_checks(t))
    t = read("test/Integration/tpp-pack-unpack.mlir")
    # this file writes some literals with parentheses: take everything between `dense<` and `> : tensor`
    lits = [nums(m.group(1)) for m in re.finditer(r"dense\s*<\s*(.*?)>\s*:\s*tensor", t, re.S)]
    add("tpp_pack_unpack_f32", "test/Integration/tpp-pack-unpack.mlir:3-91", pack1_in=lits[0], pack2_in=lits[1],
        unpack1_in=lits[3], unpack2_in=lits[4], expected_all=flat_checks(t))
    t = read("test/Integration/tiling-add.mlir")
    d = dense_blocks(t)
    add("tiling_add_f32", "test/Integration/tiling-add.mlir:14-140", A=d[0], B=d[1], expected=flat_checks(t))
    t = read("test/Integration/tiling-relu.mlir")
    add("tiling_relu_f32", "test/Integration/tiling-relu.mlir:14-105", input=dense_blocks(t)[0], expected=flat_checks(t))
    t = read("test/BF16/Integration/mlp-single-layer-bf16.mlir")
    m = re.search(r"%c256 = arith.constant (" + NUM + ")", t)
    thr = re.search(r"%threshold = arith.constant (" + NUM + ")", t)
    add("mlp_single_layer_bf16", "test/BF16/Integration/mlp-single-layer-bf16.mlir:11-55", expected_fill=float(m.group(1)),
        threshold=float(thr.group(1)))

    # round 2, fourth batch: block relayouts done as tile copies around tile gemms, the smoke-test matmul, a seeded
    # row broadcast, convolutions rewritten to matmuls over strided image windows
    t = read("test/Integration/broadcast-row-1d.mlir")
    add("broadcast_row_1d_f32_seed123", "test/Integration/broadcast-row-1d.mlir:5-26", expected=flat_checks(t))
    t = read("test/Integration/relayout-gemm.mlir")
    d = dense_blocks(t)
    add("relayout_gemm_f32", "test/Integration/relayout-gemm.mlir:30-133", A=d[0], B=d[1], expected=flat_checks(t))
    t = read("test/Integration/relayout-more-interesting.mlir")
    add("relayout_block_copy_f32", "test/Integration/relayout-more-interesting.mlir:22-105", input=dense_blocks(t)[0],
        expected_all=flat_checks(t))
    t = read("test/Integration/smoke.mlir")
    d = dense_blocks(t)
    add("smoke_matmul_f32", "test/Integration/smoke.mlir:13-59", A=d[0], B=d[1], expected=flat_checks(t))
    t = read("test/Integration/conv-to-matmul.mlir")
    add("conv_to_matmul_f32", "test/Integration/conv-to-matmul.mlir:28-152", expected_all=flat_checks(t))

    # round 2, fifth batch: seeded column broadcast, an add written into slices of a larger tensor, scalar fills of subviews
    t = read("test/Integration/broadcast-2d.mlir")
    add("broadcast_col_2d_f32_seed123", "test/Integration/broadcast-2d.mlir:7-33", expected=flat_checks(t, "COLUMNBROADCAST"))
    t = read("test/Integration/tpp-add.mlir")
    m = re.search(r"EXE-COUNT-(\d+):\s*\(([^)]*)\)", t)
    add("tpp_add_slices_f32", "test/Integration/tpp-add.mlir:23-71", repeat=int(m.group(1)), row=nums(m.group(2)))
    t = read("test/Integration/subview-on-tensor.mlir")
    add("fill_subviews_f32", "test/Integration/subview-on-tensor.mlir:11-57", expected=flat_checks(t),
        fills=[float(x) for x in re.findall(r"%cst1? = arith.constant (" + NUM + ") : f32", t)])

    with open(OUT, "w") as f:
        json.dump(g, f, indent=1, sort_keys=True)
    print(f"wrote {OUT}: {len(g)} vectors, {os.path.getsize(OUT)} bytes")


if __name__ == "__main__":
    sys.exit(main())
