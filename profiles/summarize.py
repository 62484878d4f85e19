#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small, committed
summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches.txt
  python profiles/summarize.py full gpurun_out/prof_pbs_r1.ncu-rep  > profiles/r1_pbs_v1_full.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
    "launch__block_size", "smsp__pcsamp_sample_count",
]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        name = re.sub(r"<.*", "", name)
        c = agg.setdefault(name, [0, 0.0])
        c[0] += 1
        c[1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':60s} {'launches':>8s} {'total_ms':>10s} {'avg_ms':>10s} {'share':>7s}")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:60]:60s} {n:8d} {ns/1e6:10.3f} {ns/1e6/n:10.3f} {100*ns/tot:6.1f}%")
    print(f"{'TOTAL':60s} {sum(v[0] for v in agg.values()):8d} {tot/1e6:10.3f}")


def full(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("kernel:", d.get("Kernel Name", ("", ""))[1][:90])
        for k in KEYS:
            if k in d:
                print(f"  {k:78s} {d[k][1]:>18s} {d[k][0]}")
        stalls = {k: float(v[1]) for k, v in d.items() if k.startswith("smsp__pcsamp_warps_issue_stalled_")
                  and not k.endswith("_not_issued")}
        tot = sum(stalls.values()) or 1.0
        print("  warp stall samples (share of all samples):")
        for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:12]:
            print(f"    {k.replace('smsp__pcsamp_warps_issue_stalled_', ''):28s} {100*v/tot:5.1f}%")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r)
    hdr = rows[hi]
    si, ni = hdr.index("Source"), hdr.index("Instructions Executed")
    ops = collections.Counter()
    for r in rows[hi + 1:]:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_]+(\.[A-Z0-9_]+){0,2})", r[si].strip())
        if not m:
            continue
        op = m.group(2)
        base = op.split(".")[0]
        key = op if base in ("LDS", "STS", "LDG", "STG", "LDL", "STL") else base
        ops[key] += int(r[ni] or 0)
    tot = sum(ops.values()) or 1
    print("  instruction mix (warp instructions, share):")
    for k, v in ops.most_common(22):
        print(f"    {k:16s} {v:16d} {100*v/tot:5.1f}%")


def table(path):
    """One block per captured kernel launch (a multi-kernel report): the roofline-relevant counters only."""
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        g = lambda key: d.get(key, ("", "n/a"))
        name = re.sub(r"^void ", "", g("Kernel Name")[1])
        print("kernel:", name[:110])
        print(f"  grid {g('launch__grid_size')[1]} x block {g('launch__block_size')[1]}, "
              f"{g('launch__registers_per_thread')[1]} regs, duration {g('gpu__time_duration.sum')[1]} {g('gpu__time_duration.sum')[0]}")
        print(f"  dram read {g('dram__bytes_read.sum')[1]} {g('dram__bytes_read.sum')[0]}, write {g('dram__bytes_write.sum')[1]} "
              f"{g('dram__bytes_write.sum')[0]}, dram {g('dram__throughput.avg.pct_of_peak_sustained_elapsed')[1]} % of peak, "
              f"L2 hit {g('lts__t_sector_hit_rate.pct')[1]} %, L2 {g('lts__throughput.avg.pct_of_peak_sustained_elapsed')[1]} % of peak")
        print(f"  fp64 pipe {g('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active')[1]} %, tensor pipe "
              f"{g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active')[1]} %, alu pipe "
              f"{g('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active')[1]} %, issue "
              f"{g('smsp__issue_active.avg.pct_of_peak_sustained_active')[1]} %, shared-memory wavefronts "
              f"{g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed')[1]} % of peak, "
              f"warps active {g('sm__warps_active.avg.pct_of_peak_sustained_active')[1]} %")
        stalls = {k: float(v[1]) for k, v in d.items() if k.startswith("smsp__pcsamp_warps_issue_stalled_")
                  and not k.endswith("_not_issued")}
        tot = sum(stalls.values()) or 1.0
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:4]
        print("  top stalls: " + ", ".join(f"{k.replace('smsp__pcsamp_warps_issue_stalled_', '')} {100*v/tot:.0f}%" for k, v in top))


if __name__ == "__main__":
    {"launches": launches, "full": full, "table": table}[sys.argv[1]](sys.argv[2])
