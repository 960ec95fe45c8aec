#!/usr/bin/env python
"""Turn an Nsight Compute report into the small text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/r01_step_kernel.txt

Needs `ncu` on PATH (no GPU required to read a report)."""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.per_cycle_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    name_i = hdr.index("Kernel Name")
    print("report: %s   (ncu --set full --clock-control none; cold-cache, serialised replays)" % rep)
    for r in data:
        print("\nkernel: %s" % r[name_i])
        for i, h in enumerate(hdr):
            if h in KEYS:
                print("  %-70s %s %s" % (h, r[i], units[i]))
        stalls = [(float(r[i] or 0), h) for i, h in enumerate(hdr)
                  if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
        print("  warp-cycles per issued instruction by stall reason:")
        for v, h in sorted(stalls, reverse=True)[:9]:
            print("    %6.2f  %s" % (v, h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", "")))
    src = ncu_csv(rep, "source", ("--print-source", "sass,cuda"))
    agg, cur, head = collections.Counter(), None, None
    for r in src:
        if not r:
            continue
        if r[0] == "File Path":
            cur, head = r[1].split("/")[-1], None
        elif r[0] == "Line No":
            head = r
            ie = head.index("Instructions Executed")
        elif head and cur and r[0].isdigit():
            try:
                agg[cur] += int(r[ie] or 0)
            except ValueError:
                pass
    tot = sum(agg.values())
    if tot:
        print("\nexecuted warp-instructions by source file (all profiled launches):")
        for f, v in agg.most_common():
            print("  %5.1f %%  %s" % (100.0 * v / tot, f))


if __name__ == "__main__":
    main(sys.argv[1])
