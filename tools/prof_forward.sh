#!/bin/bash
# ncu --set full capture of forward_kernel (one launch, 4096 chains of the target shape) -> gpurun_out/$1.ncu-rep
ncu --set full --clock-control none --import-source on -k regex:"forward_kernel|forward_ws_kernel" -s 3 -c 1 -f -o gpurun_out/$1 python tools/exp_time.py rf_inv_b200/librfinv_b200.so 4096 > gpurun_out/$1.log 2>&1
tail -2 gpurun_out/$1.log
