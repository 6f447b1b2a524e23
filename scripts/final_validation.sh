#!/bin/bash
# round-end validation on one B200: GPU parity tests, smoke(), the bench line (what the driver runs)
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["steps"], d["gpu_launches"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step_repetitions"],
      d["e2e"]["steps"], d["roofline"]["frac"], d["roofline"]["tensor"]["frac_of_burst_peak"], d["clocks"], d["cpu_baseline"]["value"])
PY
