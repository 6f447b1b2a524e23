#!/usr/bin/env python
"""Extract the mlir-gen invocations of the reference's benchmark configurations into a fixture.

Run in the build container (needs /root/reference):

    python tests/golden/make_bench_shapes.py

Reads benchmarks/config/{fc,matmul,omp,base}/*.json, keeps every distinct `mlir-gen` command line that describes a
matmul / fully-connected / MLP kernel of this path (flags --kernel, --bias, --relu, --float-type, --vnni, --batch,
--layers, --tiles) and writes tests/golden/reference_bench_shapes.json: one entry per distinct shape with the config
files that hold it. Nothing is computed here; tests/test_reference_bench_shapes.py replays every entry's call stream.
"""
from __future__ import annotations

import glob
import json
import os
import re
import sys

REF = os.environ.get("TPP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_bench_shapes.json")


def parse(cmd):
    def opt(name, default=None):
        m = re.search(r"--" + name + r"[= ]([^\s\"]+)", cmd)
        return m.group(1) if m else default

    layers = opt("layers")
    if layers is None:
        return None
    tiles = opt("tiles")
    return {
        "kernel": opt("kernel", "const"),
        "bias": "--bias" in cmd,
        "relu": "--relu" in cmd,
        "float_type": opt("float-type", "f32"),
        "vnni": int(opt("vnni", "0")),
        "batch": int(opt("batch", "256")),
        "layers": [int(x) for x in layers.split(",")],
        "tiles": [int(x) for x in tiles.split(",")] if tiles else None,
    }


def main():
    shapes = {}
    for sub in ("fc", "matmul", "omp", "base"):
        for path in sorted(glob.glob(os.path.join(REF, "benchmarks", "config", sub, "*.json"))):
            rel = os.path.relpath(path, REF)
            with open(path) as f:
                text = f.read()
            for m in re.finditer(r'"mlir-gen"\s*,\s*"([^"]*)"', text):
                p = parse(m.group(1))
                if p is None:
                    continue
                key = json.dumps(p, sort_keys=True)
                shapes.setdefault(key, {"shape": p, "sources": []})
                if rel not in shapes[key]["sources"]:
                    shapes[key]["sources"].append(rel)
    out = sorted(shapes.values(), key=lambda e: json.dumps(e["shape"], sort_keys=True))
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"wrote {OUT}: {len(out)} distinct mlir-gen shapes")


if __name__ == "__main__":
    sys.exit(main())
