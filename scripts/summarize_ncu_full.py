"""Summarise one `ncu --set full` capture (.ncu-rep) of the pair kernel into a text file for profiles/.

    python scripts/summarize_ncu_full.py gpurun_out/r02_pair_full.ncu-rep profiles/r02_pair_kernel_ncu_full.txt

Needs `ncu` on PATH (reading a report needs no GPU).
"""
import collections
import csv
import io
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic"]


def ncu(rep, *args):
  return subprocess.run(["ncu", "-i", rep, "--csv"] + list(args), stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main(rep, out):
  rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw"))))
  hdr, units, vals = rows[0], rows[1], rows[2]
  d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
  with open(out, "w") as f:
    f.write("ncu --set full --clock-control none --import-source on -k regex:k_edge_layer_pair -s 14 -c 1 python scripts/probe_tc.py fwd\n"
            "(one full-size launch of a middle layer, E = 400000 = 3125 tiles of 128 edges; times under ncu are not bench values)\n")
    f.write("kernel: %s\n\n" % d["Kernel Name"][0])
    for w in WANT:
      if w in d:
        f.write("%-78s %18s %s\n" % (w, d[w][0], d[w][1]))
    st = [(h.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(v.replace(",", ""))) for h, (v, u) in d.items()
          if h.startswith("smsp__pcsamp_warps_issue_stalled") and v and not h.endswith("not_issued")]
    tot = sum(x for _, x in st)
    f.write("\nwarp-state samples (all warps incl. the 4 service warps of each CTA, which only wait):\n")
    for h, x in sorted(st, key=lambda t: -t[1])[:12]:
      f.write("  %-28s %8.0f  %.3f\n" % (h, x, x / tot))
    # SASS view: samples by opcode
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--print-source", "sass"))))
    hdr = rows[1]
    si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
    ops, opi = collections.Counter(), collections.Counter()
    for r in rows[2:]:
      if len(r) <= ii:
        continue
      op = re.sub(r"^@!?U?P\w+\s+", "", r[1].strip()).split()[0].split(".")[0]
      ops[op] += int(r[si] or 0)
      opi[op] += int(r[ii] or 0)
    tot = sum(ops.values())
    f.write("\nSASS opcodes by sample count (total %d samples, %d warp instructions executed):\n" % (tot, sum(opi.values())))
    for op, n in ops.most_common(20):
      f.write("  %-12s samples %6d  %.3f   warp-instructions %10d\n" % (op, n, n / tot, opi[op]))
  print(open(out).read())


if __name__ == "__main__":
  main(sys.argv[1], sys.argv[2])
