#!/bin/bash
# ncu evidence for the path kernels (GPU box, 1 GPU), per the recipe in B200_PROFILING.md:
#   tools/gpu_profile.sh c5r c4 c2     -> gpurun_out/launches_<w>.csv, gpurun_out/prof_<w>.ncu-rep, summaries via tools/ncu_summary.py
# Each workload: the plain command first (must exit 0), then the launch list, then ONE --set full capture of the path kernel.
set -u
mkdir -p gpurun_out
for w in "$@"; do
  CMD="python bench.py --workload $w --steps 3 --warmup 3 --no-cpu --no-subs --roof-seconds 0"
  $CMD > gpurun_out/plain_$w.log 2>&1 || { echo "plain run of $w failed"; tail -5 gpurun_out/plain_$w.log; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$w.csv $CMD > gpurun_out/ncu_list_$w.log 2>&1
  echo "launch list $w rc $?"
  ncu --set full --clock-control none --import-source on -k regex:simulate_kernel -s 4 -c 2 -f -o gpurun_out/prof_$w $CMD > gpurun_out/ncu_full_$w.log 2>&1
  echo "full capture $w rc $?"
done
ls -la gpurun_out/*.ncu-rep 2>/dev/null
