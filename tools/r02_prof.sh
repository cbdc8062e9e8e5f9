#!/bin/bash
# Round-2 profile: ncu launch list of a short bench, ncu --set full of the three kernels at the bench shape, phase timing.
tag=${1:-r02a}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --pt-iters 2 > gpurun_out/bench_under_ncu_$tag.log 2>&1
grep -E "prep_kernel|forward_kernel|quadform|loglik" gpurun_out/launches_$tag.csv | awk -F'","' '{print substr($5,1,50), $(NF)}' | sort | uniq -c | sort -rn | head -20
RFINV_UPLOAD_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:"prep_kernel|forward_kernel|quadform_kernel" -s 9 -c 3 -f -o gpurun_out/prof_$tag python tools/exp_time.py rf_inv_b200/librfinv_b200.so 16384 > gpurun_out/prof_$tag.log 2>&1
tail -1 gpurun_out/prof_$tag.log
python tools/phase_timing.py --build && python tools/phase_timing.py target 2>&1 | tee gpurun_out/phases_$tag.txt
