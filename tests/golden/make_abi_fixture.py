#!/usr/bin/env python
"""Extract the reference's C-ABI of the xsmm path into a fixture (run in the build container; needs /root/reference):

    python tests/golden/make_abi_fixture.py

* the 13 prototypes of runtime/Xsmm/XsmmRunnerUtils.h:22-83 (result type, parameter list as written),
* every `call @xsmm_*(...)` FileCheck line of test/Conversion/XsmmToFunc/xsmm-to-func.mlir (the argument order the
  lowering emits: SURVEY.md 8b "exact orders are pinned by ...").
Writes tests/golden/reference_abi_calls.json; tests/test_abi_exports.py compares include/tpp_xsmm_abi.h with it.
"""
from __future__ import annotations

import json
import os
import re
import sys

REF = os.environ.get("TPP_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_abi_calls.json")


def main():
    with open(os.path.join(REF, "test/Conversion/XsmmToFunc/xsmm-to-func.mlir")) as f:
        t = f.read()
    calls = {}
    for ln, line in enumerate(t.splitlines(), 1):
        m = re.search(r"CHECK:.*call @(xsmm_\w+)\((.*)\)\s*$", line)
        if m:
            args = [a.strip() for a in m.group(2).split(",")]
            calls.setdefault(m.group(1), []).append({"line": ln, "num_args": len(args), "args": args})
    with open(os.path.join(REF, "runtime/Xsmm/XsmmRunnerUtils.h")) as f:
        h = f.read()
    hdr = {}
    for m in re.finditer(r"(\w+)\s*\n?\s*(xsmm_\w+)\(([^;]*?)\);", h, re.S):
        params = [re.sub(r"\s+", " ", p.strip()) for p in m.group(3).split(",") if p.strip()]
        hdr[m.group(2)] = {"result": m.group(1), "params": params}
    out = {"source": {"calls": "test/Conversion/XsmmToFunc/xsmm-to-func.mlir",
                      "header": "runtime/Xsmm/XsmmRunnerUtils.h:22-83"},
           "calls": calls, "header": hdr}
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(f"wrote {OUT}: {len(hdr)} prototypes, {sum(len(v) for v in calls.values())} pinned call lines")


if __name__ == "__main__":
    sys.exit(main())
