#!/bin/bash
# Run on the GPU box via gpurun: smoke, parity tests, microbench, bench.  Everything is logged under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/nvsmi.csv 2>&1
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "== tests"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
echo "== microbench"; timeout 300 python tools/microbench.py > gpurun_out/microbench.json 2> gpurun_out/microbench.err; tail -3 gpurun_out/microbench.err
for w in ${WORKLOADS:-c2}; do
  echo "== bench $w"; timeout 600 python bench.py --workload $w --steps ${STEPS:-200} --warmup 10 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; tail -3 gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json
done
