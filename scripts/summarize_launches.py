"""Summarise an ncu launch list (CSV with gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) of the denoise loop.

    python scripts/summarize_launches.py gpurun_out/r02_launches.csv profiles/r02_launches_denoise_step.txt profiles/r02_step_traffic.json

One denoise step = the launches after a k_head up to and including the next k_head.  Times under ncu are cold-cache and serialised:
only the SHARES are meaningful; the DRAM bytes are what bench.py reports as roofline.traffic (per step).
"""
import collections
import csv
import io
import json
import sys


def main(src, out_txt, out_json):
  lines = open(src).read().splitlines()
  start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
  per = collections.OrderedDict()
  for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
    per.setdefault((int(r["ID"]), r["Kernel Name"]), {})[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
  launches = [(k[0], k[1].split("(")[0].replace("void ", ""), v) for k, v in sorted(per.items())]
  heads = [i for i, l in enumerate(launches) if l[1].startswith("k_head")]
  if len(heads) < 2:
    raise SystemExit("need two k_head launches to delimit a step")
  step = launches[heads[0] + 1: heads[1] + 1]
  agg = collections.OrderedDict()
  for _, name, m in step:
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += m["gpu__time_duration.sum"]
    a[2] += m["dram__bytes_read.sum"]
    a[3] += m["dram__bytes_write.sum"]
  t_all = sum(a[1] for a in agg.values())
  rd = sum(a[2] for a in agg.values())
  wr = sum(a[3] for a in agg.values())
  with open(out_txt, "w") as f:
    f.write("one denoise step of `DFB_GRAPH_CAPTURE=0 python bench.py --steps 1 --warmup 0` under\n"
            "`ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 120 -c 80`\n"
            "(cold-cache, serialised: shares only; %d launches in the step)\n\n" % len(step))
    f.write("%-34s %4s %12s %7s %12s %12s\n" % ("kernel", "n", "time us", "share", "DRAM rd MB", "DRAM wr MB"))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      f.write("%-34s %4d %12.1f %7.3f %12.1f %12.1f\n" % (name, a[0], a[1] / 1e3, a[1] / t_all, a[2] / 1e6, a[3] / 1e6))
    f.write("%-34s %4d %12.1f %7.3f %12.1f %12.1f\n" % ("total", len(step), t_all / 1e3, 1.0, rd / 1e6, wr / 1e6))
    f.write("\nper launch, in order:\n")
    for i, name, m in step:
      f.write("%4d %-34s %9.1f us  rd %8.1f MB  wr %8.1f MB\n" % (i, name, m["gpu__time_duration.sum"] / 1e3,
                                                                m["dram__bytes_read.sum"] / 1e6, m["dram__bytes_write.sum"] / 1e6))
  pair = [a for n, a in agg.items() if "k_edge_layer_pair" in n]
  json.dump({"source": src, "launches_per_step": len(step), "dram_bytes_per_step": rd + wr, "dram_read_bytes_per_step": rd,
             "dram_write_bytes_per_step": wr, "pair_kernel_time_share": sum(a[1] for a in pair) / t_all,
             "pair_kernel_dram_bytes_per_launch": sum(a[2] + a[3] for a in pair) / max(1, sum(a[0] for a in pair))},
            open(out_json, "w"), indent=1)
  print(open(out_txt).read().split("per launch")[0])


if __name__ == "__main__":
  main(*sys.argv[1:4])
