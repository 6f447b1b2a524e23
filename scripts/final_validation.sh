#!/bin/bash
# What the driver runs at round end, in one go on a GPU box: the GPU test suite, smoke(), both bench arms at N = 1.
#   gpurun -- bash scripts/final_validation.sh
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/pytest_gpu_final.log
python __graft_entry__.py --smoke 2>&1 | tail -1 | cut -c1-400
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -2 gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().strip().splitlines()[-1])
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "strict ms", d["e2e"]["strict"]["ms_per_forward"])
x = d["extra"]
print({k: (v.get("gflops") or v.get("gbytes_per_s") or v.get("error")) if isinstance(v, dict) else v for k, v in x.items()})
l = d["latency"]
print("latency", l["ms_per_forward"], l["one_forward_per_graph"]["ms_per_forward"], l["direct_invokes"]["ms_per_forward"],
      json.dumps(l["reference_default_stream"])[:260])
for r in x["tpp_run_standin"]["runs"]:
    print(r.get("mode"), r.get("tiles"), r.get("seconds_per_iteration"), r.get("kernel"))
PY
python bench.py --impl reference --steps 5 --warmup 2 2>/dev/null | tail -1 | cut -c1-200
