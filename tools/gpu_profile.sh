#!/bin/bash
# ncu evidence for one workload (run under gpurun, one GPU): launch list + full capture of the fused path kernel.
W=${1:-c2}
CMD="python bench.py --workload $W --steps ${NCU_STEPS:-3} --warmup 3 --roof-seconds 0 --no-cpu"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$W.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$W.csv $CMD > gpurun_out/ncu_list_$W.log 2>&1
$CMD > gpurun_out/plain2_$W.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:simulate_kernel -s 4 -c 2 -f -o gpurun_out/prof_$W $CMD > gpurun_out/ncu_full_$W.log 2>&1
tail -3 gpurun_out/ncu_full_$W.log
ls -la gpurun_out/
