#!/usr/bin/env python
"""Turn an Nsight Compute report into the small text summary committed under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/r02_step_kernels_ncu.txt
    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep --traffic iiwa 65536 profiles/r02_step_kernels_ncu.txt

The second form also (re)writes the workload's entry of profiles/dram_traffic.json — the DRAM bytes per step
(dram__bytes_read.sum + dram__bytes_write.sum, summed over the kernels of one step) that bench.py reports as
roofline.traffic — straight from the report, so the number in the bench line is the measured one.

Needs `ncu` on PATH (no GPU required to read a report)."""
import json
import os
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


def traffic(rep, workload, batch, source):
    rows = ncu_csv(rep, "raw")
    hdr, data = rows[0], rows[2:]
    name_i, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    units = rows[1]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_kernel = collections.OrderedDict()
    for r in data:                       # first profiled launch of every kernel = one step
        k = r[name_i].split("(")[0]
        if k not in per_kernel:
            per_kernel[k] = (float(r[rd]) * scale.get(units[rd], 1.0), float(r[wr]) * scale.get(units[wr], 1.0))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dram_traffic.json")
    try:
        table = json.load(open(path))
        if "bytes_per_launch" in table:  # round-1 layout (one hand-filled entry)
            table = {}
    except Exception:
        table = {}
    table[workload] = dict(batch=int(batch), bytes_per_launch=int(sum(a + b for a, b in per_kernel.values())),
                           kernels={k: dict(dram_bytes_read=int(a), dram_bytes_write=int(b)) for k, (a, b) in per_kernel.items()},
                           source="%s (ncu --set full --clock-control none of `bench.py`, first profiled launch of each "
                                  "kernel of one step; written by profiles/summarize_ncu.py)" % source)
    json.dump(table, open(path, "w"), indent=1)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[2] == "--traffic":
        traffic(sys.argv[1], sys.argv[3], sys.argv[4], sys.argv[5])
    else:
        main(sys.argv[1])
