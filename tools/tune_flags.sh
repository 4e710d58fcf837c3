#!/bin/bash
# Compiler-flag sweep of the path kernels on the GPU box (rebuilds with MCB200_NVCC_FLAGS each time).
IFS=';'
for flags in ${FLAGSETS:-;-Xptxas -O2;-Xptxas --allow-expensive-optimizations=true;-Xptxas -O1}; do
  unset IFS
  MCB200_NVCC_FLAGS="$flags" python monte-carlo-options-simd_b200/build.py --force > /dev/null 2>&1 || echo "build failed: $flags"
  for w in ${WORKLOADS:-c2 c5r c4}; do
    python bench.py --workload $w --steps $([ $w = c2 -o $w = c3 ] && echo 100 || echo 3) --warmup 3 --no-cpu --roof-seconds 0.3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('flags [$flags] $w frac %.4f kernel_ms %.4f'%(d['roofline']['frac'], d['roofline']['kernel_ms']))"
  done
  IFS=';'
done
unset IFS
python monte-carlo-options-simd_b200/build.py --force > /dev/null 2>&1
