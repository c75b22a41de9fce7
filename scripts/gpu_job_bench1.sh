#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_n1.log").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "frac", d["roofline"]["frac"], "kernel", d["roofline"]["kernel"]["ms_per_launch"], d["roofline"]["kernel"]["frac"], "share", d["roofline"]["kernel"]["share_of_step"], "launches/step", d["launches_per_denoise_step"], d["clocks"])
PY
