#!/bin/bash
# final tree: full GPU suite, headline bench (N=1), reference arm, other configs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1
echo "all rc=$?" >> gpurun_out/t_all.log
tail -n 4 gpurun_out/t_all.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.log 2> gpurun_out/bench_n1.err
tail -n 1 gpurun_out/bench_n1.log | cut -c1-400
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
tail -n 1 gpurun_out/bench_ref.log | cut -c1-600
: > gpurun_out/bench_cfg.jsonl
for c in B1 C1 C3 C4 C5; do
  timeout 600 python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline 2>> gpurun_out/bench_cfg.err | tail -n 1 >> gpurun_out/bench_cfg.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/bench_cfg.jsonl"):
    try:
        d = json.loads(l)
        print(d["metric"][:60], "value", round(d["value"], 2), "frac", round(d["roofline"]["frac"], 3), "launches/step", d["launches_per_denoise_step"])
    except Exception as e:
        print("bad line", e, l[:200])
PY
