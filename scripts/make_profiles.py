#!/usr/bin/env python
"""Turn the ncu outputs in gpurun_out/ into the tracked summaries under profiles/ (read here on CPU).
usage: scripts/make_profiles.py [config] [tag]"""
import collections, csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cfg = sys.argv[1] if len(sys.argv) > 1 else "c3"
tag = sys.argv[2] if len(sys.argv) > 2 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
os.makedirs(P, exist_ok=True)


def rows_of(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            return r, rows[i + 1:]
    raise SystemExit("no header in " + path)


# 1. launch list -> per-kernel shares of one step
h, body = rows_of(os.path.join(G, "launches_%s_%s.csv" % (cfg, tag)))
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
seq = [(r[ki], float(r[vi].replace(",", ""))) for r in body if len(r) > vi]
# last step = everything after the last dda_codes-preceding fill of the last reset: take the last 1/4 of the run
names = [n for n, _ in seq]
last_reset = max(i for i, n in enumerate(names) if n.startswith("dda_codes_kernel") and (i == 0 or not names[i - 1].startswith(("void simscore3", "planemap3", "void simmap"))) and
                 not any(m.startswith("dda_codes_kernel") for m in names[max(0, i - 3):i]))
step = seq[last_reset:]
agg = collections.OrderedDict()
for n, v in step:
    key = n.split("(")[0].replace("void ", "")
    if key.startswith("at::") or key.startswith("at_cuda") or key.startswith("<unnamed>"):
        key = "torch helper kernels"
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, "%s_launches_%s_summary.txt" % (tag, cfg)), "w") as f:
    f.write("ncu --metrics gpu__time_duration.sum --clock-control none, one bench step (%s): %d launches, %.3f ms of kernel time\n"
            "(per-launch times are cold-cache and serialised: compare SHARES with bench.py's stages_ms, not absolutes)\n\n" % (cfg, len(step), tot / 1e6))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-42s launches %4d  total %9.3f ms  %5.1f %%\n" % (k, a[0], a[1] / 1e6, 100 * a[1] / tot))
os.replace(os.path.join(G, "launches_%s_%s.csv" % (cfg, tag)), os.path.join(P, "%s_launches_%s.csv" % (tag, cfg))) if False else None
import shutil
shutil.copy(os.path.join(G, "launches_%s_%s.csv" % (cfg, tag)), os.path.join(P, "%s_launches_%s.csv" % (tag, cfg)))

# 2. DRAM traffic of one whole (non-first) sweep
h, body = rows_of(os.path.join(G, "bp_sweep_dram_%s_%s.csv" % (cfg, tag)))
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
launches = collections.OrderedDict()
for r in body:
    if len(r) <= vi:
        continue
    d = launches.setdefault(r[ii], {"name": r[ki]})
    d[r[mi]] = float(r[vi].replace(",", ""))
L = list(launches.values())
nonfirst = [d for d in L if ", 0>" in d["name"] or "false" in d["name"]]
per_sweep = len(set(d["name"] for d in nonfirst))
sweep = nonfirst[:per_sweep]
rd = sum(d["dram__bytes_read.sum"] for d in sweep)
wr = sum(d["dram__bytes_write.sum"] for d in sweep)
ns = sum(d["gpu__time_duration.sum"] for d in sweep)
traffic = {"bp_kernel_dram_bytes_per_launch": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr,
           "kernel_ns_under_ncu": ns, "launches_per_sweep": per_sweep,
           "what": "sum over the %d length-class launches of ONE non-first BP sweep (bp4_kernel<NCH, false>), config %s; "
                   "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum" % (per_sweep, cfg),
           "per_class": [{"kernel": d["name"], "dram_read": d["dram__bytes_read.sum"], "dram_write": d["dram__bytes_write.sum"],
                          "ns": d["gpu__time_duration.sum"]} for d in sweep]}
json.dump(traffic, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
shutil.copy(os.path.join(G, "bp_sweep_dram_%s_%s.csv" % (cfg, tag)), os.path.join(P, "%s_bp_sweep_dram_%s.csv" % (tag, cfg)))

# 3. full-set summaries
reps = [os.path.join(G, f) for f in ["prof_bp4_%s_%s.ncu-rep" % (cfg, tag)] +
        ["prof_%s_%s_%s.ncu-rep" % (k, cfg, tag) for k in ("simscore3_kernel", "bp4_first_mapped_kernel", "planemap3_kernel", "depth3_kernel", "dda_codes_kernel")]
        if os.path.exists(os.path.join(G, f))]
out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py")] + reps, capture_output=True, text=True).stdout
open(os.path.join(P, "%s_ncu_full_%s.txt" % (tag, cfg)), "w").write(out)
print(open(os.path.join(P, "%s_launches_%s_summary.txt" % (tag, cfg))).read())
print(json.dumps({k: v for k, v in traffic.items() if k != "per_class"}, indent=1))
