#!/usr/bin/env python
"""Turns the ncu captures of tools/final_measure.sh (gpurun_out/launches_<T>.csv, gpurun_out/prof_<T>.ncu-rep)
into the tracked summaries under profiles/: launch-list shares, the --set full table of the top kernels and
r01_traffic.json (DRAM bytes per sample launch per bench.py span, used for roofline.traffic).
usage: python tools/summarize_profiles.py r1c r01c"""
import collections
import csv
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T = sys.argv[1] if len(sys.argv) > 1 else "r1c"
OUT = sys.argv[2] if len(sys.argv) > 2 else "r01c"
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

SPAN = {"fq_count_newlines": "fq_index", "scan_u32_to_u64": "fq_index", "fq_index_lines": "fq_index", "fq_cta_pos": "fq_index", "s1_superk": "s1_superk", "s1_superk_v5": "s1_superk",
        "hash_hist_roll_kernel": "hash_hist", "hash_hist_kernel": "hash_hist", "hash_compact_kernel": "hash_emit",
        "hash_scan_kernel": "hash_emit", "hash_copy_kernel": "hash_emit", "merge_emit_kernel": "merge", "merge_solid_kernel": "merge"}


def short(name):
    m = re.match(r"(?:void )?([A-Za-z0-9_]+)", name)
    return m.group(1) if m else name


# ---- launch list
rows = [r for r in csv.reader(l for l in open(os.path.join(G, f"launches_{T}.csv")) if l.startswith('"'))]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    tot[short(r[ki])] += float(r[vi].replace(",", "")) / 1e3
    cnt[short(r[ki])] += 1
hot = {k: v for k, v in tot.items() if k in SPAN}
s_hot = sum(hot.values())
with open(os.path.join(P, f"{OUT}_launches_summary.md"), "w") as f:
    f.write(f"# Round 1 (final state) -- ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n")
    f.write("Command: `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_%s.csv python bench.py "
            "--samples 4 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --lanes 1`\n(4 samples x 1M reads x 150 nt, k=31, hash:bf:bin, P=64; "
            "serialised, cold cache: compare SHARES).\n\n| kernel | bench.py span | launches | total us | share of hot-path kernels |\n|---|---|---|---|---|\n" % T)
    for k, v in sorted(tot.items(), key=lambda t: -t[1]):
        sh = f"{100 * v / s_hot:.1f}%" if k in SPAN else "(not on the timed path)"
        f.write(f"| {k} | {SPAN.get(k, '-')} | {cnt[k]} | {v:.1f} | {sh} |\n")
    span_tot = collections.Counter()
    for k, v in hot.items():
        span_tot[SPAN[k]] += v
    f.write("\nPer bench.py span: " + ", ".join(f"{k} {100 * v / s_hot:.1f}%" for k, v in sorted(span_tot.items(), key=lambda t: -t[1])) + "\n")
import shutil
shutil.copy(os.path.join(G, f"launches_{T}.csv"), os.path.join(P, f"{OUT}_launches.csv"))

# ---- --set full
raw = subprocess.run(["ncu", "-i", os.path.join(G, f"prof_{T}.ncu-rep"), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__grid_size"]
idx = [hdr.index(w) for w in want]
ki = hdr.index("Kernel Name")


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def to_ms(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(u, 1)


spans = collections.defaultdict(lambda: {"dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "time_ms": 0.0, "kernels": []})
seen = set()
with open(os.path.join(P, f"{OUT}_ncu_full_summary.md"), "w") as f:
    f.write("# Round 1 (final state) -- ncu --set full, hot-path kernels of ONE sample launch (1M reads x 150 nt, 1.2e8 k-mers, k=31, hash keys, P=64, bloom 2e8)\n\n")
    f.write("Command: `ncu --set full --clock-control none --import-source on -k regex:\"s1_superk|hash_hist_roll|hash_compact|hash_copy|hash_scan|"
            "fq_index_lines|fq_count_newlines|fq_cta_pos\" -s 7 -c 7 -o gpurun_out/prof_%s python bench.py --samples 2 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --lanes 1`\n\n" % T)
    f.write("| kernel | " + " | ".join(w.replace(".avg.pct_of_peak_sustained", ".pct") for w in want) + " |\n|" + "---|" * (len(want) + 1) + "\n")
    # ncu reports a per-row unit in the metric-unit row only; values with mixed units are normalised below
    for r in rows[2:]:
        name = short(r[ki])
        if name in seen:
            continue
        seen.add(name)
        vals = []
        for w, i in zip(want, idx):
            v, u = r[i], units[i]
            if w == "gpu__time_duration.sum":
                vals.append(f"{to_ms(v, u) * 1e3:.1f} us")
            elif w.startswith("dram__bytes"):
                vals.append(f"{to_bytes(v, u) / 1e6:.1f} MB")
            else:
                vals.append(v)
        f.write(f"| {name} | " + " | ".join(vals) + " |\n")
        sp = SPAN.get(name)
        if sp:
            spans[sp]["dram_read_bytes"] += to_bytes(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
            spans[sp]["dram_write_bytes"] += to_bytes(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
            spans[sp]["time_ms"] += to_ms(r[hdr.index("gpu__time_duration.sum")], units[hdr.index("gpu__time_duration.sum")])
            spans[sp]["kernels"].append(name)
json.dump({"source": f"ncu --set full --clock-control none, see profiles/{OUT}_ncu_full_summary.md (one sample launch = 1M reads x 150 nt, k=31, hash keys, P=64, bloom 2e8); "
                     "per bench.py span = sum over the span's kernels", "spans": spans},
          open(os.path.join(P, "r01_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{OUT}_ncu_full_summary.md")).read())
print(open(os.path.join(P, f"{OUT}_launches_summary.md")).read())
