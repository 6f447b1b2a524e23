"""patches/*.patch must apply to the reference tree and only call entry points the ABI header
declares. CPU-only; skipped where the reference checkout is not present (the GPU box)."""
import glob
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PATCHES = sorted(glob.glob(os.path.join(ROOT, "patches", "*.patch")))


def test_patch_set_is_present():
    assert len(PATCHES) >= 5


@pytest.mark.skipif(not os.path.isdir(REF) or shutil.which("patch") is None, reason="needs the reference checkout and patch(1)")
@pytest.mark.parametrize("patch", PATCHES, ids=[os.path.basename(p) for p in PATCHES])
def test_patch_applies_to_the_reference(patch, tmp_path):
    files = re.findall(r"^\+\+\+ b/(\S+)", open(patch).read(), flags=re.M)
    assert files
    for f in files:   # copy only the touched files: the reference itself is read-only
        dst = tmp_path / f
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copy(os.path.join(REF, f), dst)
    r = subprocess.run(["patch", "-p1", "--dry-run", "-s", "-i", patch], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_patches_only_call_declared_abi_symbols():
    header = open(os.path.join(ROOT, "include", "tpp_xsmm_abi.h")).read()
    declared = set(re.findall(r"\b(xsmm_\w+|perf_\w+_timer|libxsmm_cpuid_dot_pack_factor)\s*\(", header))
    used = set()
    for p in PATCHES:
        for line in open(p):
            if line.startswith("+") and not line.startswith("+++"):
                used |= set(re.findall(r'"(xsmm_cuda_\w+)"', line))
                used |= set(re.findall(r"\b(libxsmm_cpuid_dot_pack_factor)\s*\(", line))
    assert used, "the patches are expected to call the CUDA extensions"
    assert used <= declared, used - declared
