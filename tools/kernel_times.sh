#!/bin/bash
# per-kernel device time and FP64 pipe utilisation of one evaluation (target shape, 16384 chains) under ncu
so=${1:-rf_inv_b200/librfinv_b200.so}
ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
  -k regex:"prep_kernel|forward_kernel|forward_ws_kernel|quadform" -s 6 -c 3 --csv python tools/exp_time.py $so ${2:-16384} 2>/dev/null \
  | grep -E '^"[0-9]' | awk -F'","' '{split($5,a,"("); printf "%-60s %-70s %s\n", substr(a[1],1,60), $(NF-2), $NF}'
