#!/bin/bash
# per-kernel device times (ncu, cold cache) + CUDA-event times of several builds of the library: tools/r02_ab.sh tag so1 so2 ...
tag=$1; shift
mkdir -p gpurun_out
for so in "$@"; do
  echo "== $so"
  ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none \
    -k regex:"prep_kernel|forward_kernel|quadform" -s 6 -c 3 --csv python tools/exp_time.py $so 16384 2>/dev/null \
    | grep -E '^"[0-9]' | awk -F'","' '{split($5,a,"("); printf "%-40s %-60s %s\n", substr(a[1],1,40), $(NF-2), $NF}'
  python tools/exp_time.py $so 16384 2>&1 | tail -1
done | tee gpurun_out/ab_$tag.txt
