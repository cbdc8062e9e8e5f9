#!/bin/bash
# Round-2 GPU check: parity tests, A/B kernel times of two library builds, one bench line.
#   tools/r02_run.sh <tag> [old.so]
tag=${1:-r02a}
old=${2:-rf_inv_b200/librfinv_b200_r01.so}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_$tag.log
tail -5 gpurun_out/pytest_$tag.log
for so in $old rf_inv_b200/librfinv_b200.so; do
  [ -f $so ] && timeout 300 python tools/exp_time.py $so 16384 2>&1 | tail -1
done | tee gpurun_out/ab_$tag.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 400 gpurun_out/bench_$tag.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$tag.json"))
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["roofline"]["kernel_ms"], d["roofline"]["whole_step_frac"], d["pt"] and d["pt"]["iters_per_s"])
PY
