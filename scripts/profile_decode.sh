timeout 400 python tests/tools/bench_decode.py --sizes 2000,5000,10000 --two_opt_iterations 300 --out gpurun_out/decode_bench_large.jsonl 2>&1 | tail -4
timeout 200 ncu --set full --clock-control none -k regex:k_twoopt_eval -s 20 -c 1 -o gpurun_out/twoopt_eval python tests/tools/bench_decode.py --sizes 5000 --two_opt_iterations 40 --out /tmp/x.jsonl > gpurun_out/ncu_twoopt.log 2>&1
ncu -i gpurun_out/twoopt_eval.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; r=rows[-1]
want=['gpu__time_duration.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum']
for w in want:
  for i,n in enumerate(h):
    if n==w: print(w, r[i], rows[1][i])
" > gpurun_out/twoopt_eval_summary.txt 2>&1; cat gpurun_out/twoopt_eval_summary.txt
