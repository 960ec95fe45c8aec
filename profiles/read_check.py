"""Print kernel durations (ncu launch list) and the bench line of a profiles/gpu_check.sh pass: python profiles/read_check.py TAG"""
import collections
import csv
import json
import sys

tag = sys.argv[1]
try:
    rows = list(csv.reader(open("gpurun_out/%s_launches.csv" % tag)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    kn, mv = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.defaultdict(list)
    for r in rows[hdr + 2:]:
        if len(r) > mv:
            try:
                agg[r[kn][:64]].append(float(r[mv].replace(",", "")))
            except ValueError:
                pass
    for k, v in agg.items():
        print("%-66s n=%3d mean %.1f us min %.1f max %.1f" % (k, len(v), sum(v) / len(v) / 1e3, min(v) / 1e3, max(v) / 1e3))
except Exception as exc:
    print("no launch list:", exc)
try:
    l = json.load(open("gpurun_out/%s_bench.json" % tag))
    c = l["config"]
    print("bench: %.2f us/step (min %.2f), value %.3g, canonical-basis kernel %s us, redone %s, e2e %.3g staged %.3g, frac %.4f"
          % (l["ms_per_step"] * 1e3, c["min_ms_per_step"] * 1e3, l["value"],
             None if c.get("canonical_basis_kernel_ms_per_step") is None else round(c["canonical_basis_kernel_ms_per_step"] * 1e3, 2),
             c.get("share_redone_with_lapack_basis"), l["e2e"]["value"], l["e2e"]["staged_value"], l["roofline"]["frac"]))
except Exception as exc:
    print("no bench line:", exc)
