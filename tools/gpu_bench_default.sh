#!/bin/bash
# usage (under gpurun): bash tools/gpu_bench_default.sh <tag>  -- the driver's default bench call, timed, with a summary
TAG=$1
mkdir -p gpurun_out
t0=$(date +%s)
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench.py wall $(( $(date +%s) - t0 )) s"
tail -c 400 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("c3 ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e ms", d["e2e"]["ms_per_step"], "cpu", d.get("cpu_baseline", {}).get("value"))
for k, v in d["workloads"].items():
    print(k, {x: v.get(x) for x in ("ms_per_step", "roofline_frac", "fp64_frac", "error", "wall_s")})
print(d["e2e_dropin"])
PY
