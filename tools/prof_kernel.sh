#!/bin/bash
# ncu --set full capture of one launch of kernel regex $1 (target shape, $3 chains, default 4096) -> gpurun_out/$2.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"$1" -s 3 -c 1 -f -o gpurun_out/$2 python tools/exp_time.py rf_inv_b200/librfinv_b200.so ${3:-4096} > gpurun_out/$2.log 2>&1
tail -2 gpurun_out/$2.log
