#!/bin/bash
# per-kernel device times of the PT iteration (ncu, serialised): tools/pt_launches.sh tag
tag=${1:-r02}
mkdir -p gpurun_out
RFINV_PT_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches_pt_$tag.csv python tools/pt_time.py 40 > gpurun_out/pt_under_ncu_$tag.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_pt_$tag.csv")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]; iN = h.index("Kernel Name"); iV = h.index("Metric Value")
t = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    try: t[r[iN].split("(")[0][-44:]].append(float(r[iV].replace(",", "")))
    except Exception: pass
tot = 0
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    print("%-46s n=%3d mean %8.1f us  total %9.1f" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3)); tot += sum(v)
n_it = max(len(v) for v in t.values())
print("sum per iteration (us):", tot / 1e3 / n_it)
PY
