#!/bin/bash
# Round evidence: bench line, ncu launch list of the same command, one ncu --set full capture per kernel at the bench shape.
tag=${1:-r01x}
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 600 gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --pt-iters 2 > gpurun_out/bench_under_ncu_$tag.log 2>&1
RFINV_UPLOAD_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:"prep_kernel|forward_kernel|quadform_kernel" -s 9 -c 3 -f -o gpurun_out/prof_$tag python tools/exp_time.py rf_inv_b200/librfinv_b200.so 16384 > gpurun_out/prof_$tag.log 2>&1
tail -1 gpurun_out/prof_$tag.log
cat gpurun_out/bench_$tag.json
