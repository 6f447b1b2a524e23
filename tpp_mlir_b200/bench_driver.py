#!/usr/bin/env python
"""The reference's benchmark driver for this backend.

`benchmarks/driver.py -c <config.json>` of the reference (benchmarks/driver.py:417-515) reads a JSON list of
benchmarks, each a set of named runs, executes every run the machine supports and prints

    Benchmark: <name>
    <run name, 28 wide>: <number> gflops

This module keeps that contract - same JSON format, same `-c/--config` (comma-separated files), `-n`, `--build`,
`--seed`, `-v/-q`, `--ignore-errors`, same output lines (benchmarks/harness/controller.py:314-318: `%9.3f gflops`) - for
the runs that are this path: `"type": "IR-GEN"` entries whose generator is `mlir-gen` with a matmul / fully-connected /
MLP kernel. Their flags go, as they stand, to `tpp_run_standin --mlir-gen "<flags>"` (csrc/harness/tpp_run_standin.cpp:
the program `mlir-gen ... | tpp-run -n N` would run, on the C-ABI of libtpp_xsmm_runner_utils.so), by default in the
`graph` mode of the patched runner (patches/0005). Runs of other types need the reference's own toolchain - `MLIR`
(tpp-opt / tpp-run on an .mlir file), `XSMM-DNN` (the libxsmm-dnn binary), `GENERIC` - and are reported as skipped.

    python -m tpp_mlir_b200.bench_driver -c /path/to/tpp-mlir/benchmarks/config/omp/mlir-bf16.json [-n 100]
                                         [--mode strict|device|lazy|graph] [--build DIR] [--ignore-extensions] [--json]

`extensions` lists regular expressions on the host's CPU feature flags (benchmarks/driver.py:73-101). They select CPU
ISA variants of one and the same kernel; like the reference this driver runs an entry when one of them matches the
host, which keeps one run per shape on a given box. `--ignore-extensions` runs every entry.
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def cpu_flags():
    """feature flags of the host CPU, `flags` (x86) or `Features` (Arm) of /proc/cpuinfo"""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith(("flags", "Features")):
                    return line.split(":", 1)[1].split()
    except OSError:
        pass
    return []


def supported(extensions, flags):
    """an empty list means any machine; otherwise one of the expressions has to match one flag completely"""
    return not extensions or any(re.fullmatch(ext, flag) for ext in extensions for flag in flags)


class Run:
    """one named run of a benchmark: what to execute, and what came out"""

    def __init__(self, name, spec):
        self.name = name
        self.kind = spec.get("type", "")
        self.benchmark = spec.get("benchmark")
        self.environment = dict(spec.get("environment") or {})
        flags = spec.get("flags") or []
        self.flags = list(flags) if isinstance(flags, list) else []
        self.extensions = list(spec.get("extensions") or [])
        self.stdout = ""
        self.stderr = ""
        self.gflops = None
        self.skipped = None      # reason, if this backend does not run the entry
        self.row = None          # the stand-in's JSON line

    def generator_flags(self):
        """the mlir-gen flags of an IR-GEN entry (`"benchmark": ["mlir-gen", "--kernel=... --batch=..."]`), or None"""
        b = self.benchmark
        if self.kind != "IR-GEN" or not isinstance(b, list) or not b or os.path.basename(b[0]) != "mlir-gen":
            return None
        return " ".join(b[1:]).strip()

    def iterations(self, forced):
        if forced:
            return str(int(forced))
        if "-n" in self.flags and self.flags.index("-n") + 1 < len(self.flags):
            return str(int(self.flags[self.flags.index("-n") + 1]))
        return "100"

    def command(self, exe, args):
        cmd = [exe, "--mlir-gen", self.generator_flags(), "-n", self.iterations(args.n), "--mode", args.mode]
        if args.seed:
            cmd += ["--seed", str(int(args.seed))]
        return cmd

    def execute(self, exe, args):
        gen = self.generator_flags()
        if gen is None:
            self.skipped = {"MLIR": "needs tpp-opt / tpp-run (the compiler is not part of this backend)",
                            "XSMM-DNN": "the libxsmm-dnn CPU binary",
                            "GENERIC": "an arbitrary command line"}.get(self.kind, f"run type '{self.kind}'")
            return True
        env = dict(os.environ)
        env.update({k: str(v) for k, v in self.environment.items()})
        try:
            r = subprocess.run(self.command(exe, args), capture_output=True, text=True, env=env, timeout=args.timeout)
        except (OSError, subprocess.TimeoutExpired) as e:
            self.stderr = repr(e)
            return False
        self.stderr = r.stderr
        if r.returncode != 0:
            return False
        try:
            self.row = json.loads(r.stdout.strip().splitlines()[-1])
            self.gflops = float(self.row["gflops"])
        except (ValueError, KeyError, IndexError):
            self.stderr += "\ncannot read the stand-in's result line: " + r.stdout[-200:]
            return False
        self.stdout = f"{self.gflops:9.3f} gflops"     # controller.py:316
        return True


def read_configs(paths, flags, ignore_extensions, log):
    """[(benchmark name, [Run, ...]), ...] in file order; raises SyntaxError like the reference on a bad file"""
    out = []
    for path in paths.split(","):
        if not os.path.exists(path):
            raise SyntaxError(f"Cannot find JSON config '{path}'")
        with open(path) as f:
            cfg = json.load(f)
        for entry in cfg:
            if not isinstance(entry, dict) or len(entry) != 1:
                raise SyntaxError("List of dict with a single element expected")
            (name, runs), = entry.items()
            selected = []
            for key, spec in runs.items():
                if not ignore_extensions and not supported(spec.get("extensions") or [], flags):
                    log(f"Skipping {key} as its extensions {spec.get('extensions')} are not supported")
                    continue
                selected.append(Run(key, spec))
            out.append((name, selected))
    return out


def main(argv=None):
    ap = argparse.ArgumentParser(description="TPP-MLIR benchmark driver, B200 xsmm backend")
    ap.add_argument("-c", "--config", type=str, default="benchmarks.json", help="JSON file(s) containing benchmark configuration")
    ap.add_argument("-n", type=str, default="", help="Force number of iterations on all benchmarks")
    ap.add_argument("--build", type=str, default="", help="Directory that holds tpp_run_standin (default: tpp_mlir_b200/lib)")
    ap.add_argument("-v", "--verbose", action="count", default=0)
    ap.add_argument("-q", "--quiet", action="count", default=0)
    ap.add_argument("--ignore-errors", action="count", default=0, help="Ignore errors and only show the results that work")
    ap.add_argument("--seed", type=str, help="Random seed")
    ap.add_argument("--mode", default="graph", choices=["strict", "device", "lazy", "graph"],
                    help="operand residency / issue mode of the stand-in (see tpp_run_standin.cpp)")
    ap.add_argument("--ignore-extensions", action="store_true", help="run every entry whatever CPU features it names")
    ap.add_argument("--json", action="store_true", help="additionally print one JSON line per run that was executed")
    ap.add_argument("--timeout", type=float, default=600.0)
    args = ap.parse_args(argv)

    def log(msg):
        if args.verbose > args.quiet:
            print(msg, file=sys.stderr)

    exe = os.path.join(args.build or os.path.join(HERE, "lib"), "tpp_run_standin")
    if not os.path.exists(exe):
        if args.build:
            print(f"no tpp_run_standin in '{args.build}'", file=sys.stderr)
            return 1
        from . import _build

        exe = _build.build_standin()
    try:
        benchmarks = read_configs(args.config, cpu_flags(), args.ignore_extensions, log)
    except (SyntaxError, ValueError) as e:
        print(f"Error finding benchmarks: {e}", file=sys.stderr)
        ap.print_help()
        return 1
    for name, runs in benchmarks:
        for run in runs:
            if not run.execute(exe, args) and not args.ignore_errors:
                print(f"Error executing the benchmarks: {name} / {run.name}: {run.stderr.strip()}", file=sys.stderr)
                return 1
    # benchmarks/driver.py:498-515
    for name, runs in benchmarks:
        print(f"Benchmark: {name}")
        for run in runs:
            if run.skipped:
                log(f"{run.name}: skipped ({run.skipped})")
                continue
            if not run.stdout:
                print(f"Benchmark {name}, run {run.name} produced no output, can't verify results: {run.stderr.strip()}",
                      file=sys.stderr)
                if not args.ignore_errors:
                    return 1
                continue
            print(f"{run.name:28}: {run.stdout}")
            if args.json:
                print(json.dumps({"benchmark": name, "run": run.name, "mlir_gen": run.generator_flags(), **run.row}))
        print("")
    return 0


if __name__ == "__main__":
    sys.exit(main())
