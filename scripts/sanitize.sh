#!/bin/bash
# compute-sanitizer over the parity suite (race / memory checking of the hot path).
#   scripts/sanitize.sh [memcheck|racecheck|synccheck|initcheck] [pytest -k expression]
# Runs on a GPU box (gpurun); writes gpurun_out/sanitize_<tool>.log and prints the summary lines. The default test
# selection covers every kernel family once (per-layer tcgen05 kernels incl. split-K, the chain kernels, the
# pair-per-chain kernel with blocked / VNNI-2 operands, eltwise / transpose / VNNI pack, batched tile moves) at sizes
# the sanitizer finishes in minutes.
TOOL=${1:-memcheck}
SEL=${2:-"test_fused_brgemm_bf16 or test_brgemm_bf16_tensor_core_path or regrouped or combined or reference_default or vnni2_pack or vnni4_pack or transpose_bit_exact or tiled_pack or captured_chain or many_captured or binary_vs_oracle or lone_blocked or few_blocked or unrolled_blocked"}
mkdir -p gpurun_out
LOG=gpurun_out/sanitize_${TOOL}.log
timeout 1500 compute-sanitizer --tool ${TOOL} --print-limit 20 --error-exitcode 99 \
  python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "${SEL}" > ${LOG} 2>&1
RC=$?
echo "compute-sanitizer ${TOOL}: exit code ${RC}"
grep -E "ERROR SUMMARY|passed|failed|error" ${LOG} | tail -5
exit ${RC}
