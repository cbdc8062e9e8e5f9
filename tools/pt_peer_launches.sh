#!/bin/bash
# per-kernel device times of the PT iteration in the 2-process peer-memory mode: process 0 runs under ncu (one pass, no replay),
# process 1 plain.  tools/pt_peer_launches.sh tag
tag=${1:-r02}
mkdir -p gpurun_out
cat > /tmp/peer_wrap.sh <<'W'
#!/bin/bash
if [ "$LOCAL_RANK" = "0" ]; then
  exec ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 100 --csv --log-file gpurun_out/launches_pt_peer_$TAG.csv python tools/pt_exchange_time.py 60
else
  exec python tools/pt_exchange_time.py 60
fi
W
chmod +x /tmp/peer_wrap.sh
TAG=$tag RFINV_PT_GRAPH=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 --no-python /tmp/peer_wrap.sh > gpurun_out/pt_peer_under_ncu_$tag.log 2>&1
python - <<PY
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_pt_peer_$tag.csv")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]; iN = h.index("Kernel Name"); iV = h.index("Metric Value")
t = collections.defaultdict(list)
for r in rows[hdr + 1:]:
    try: t[r[iN].split("(")[0][-44:]].append(float(r[iV].replace(",", "")))
    except Exception: pass
for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
    print("%-46s n=%3d mean %8.1f us  total %9.1f" % (k, len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3))
PY
