#!/bin/bash
# Sweep the two compile-time knobs of the path kernels on the GPU box (rebuilds with MCB200_NVCC_FLAGS each time):
#   KERNEL=PRICE|GREEKS  CONFIGS="unroll:ctas ..."  WORKLOADS="c2 c5r" (PRICE) or "c4" (GREEKS)
for cfg in ${CONFIGS:-2:4 1:4 4:4 3:4 2:3 2:2}; do
  u=${cfg%%:*}; c=${cfg##*:}
  MCB200_NVCC_FLAGS="-DMCB_${KERNEL:-PRICE}_UNROLL=$u -DMCB_${KERNEL:-PRICE}_CTAS=$c" python monte-carlo-options-simd_b200/build.py --force > /dev/null 2>&1 || echo "build failed for $cfg"
  for w in ${WORKLOADS:-c2 c5r c4}; do
    python bench.py --workload $w --steps $([ $w = c2 -o $w = c3 ] && echo 100 || echo 3) --warmup 3 --no-cpu --roof-seconds 0.3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('unroll $u ctas $c $w frac %.4f kernel_ms %.4f'%(d['roofline']['frac'], d['roofline']['kernel_ms']))"
  done
done
python monte-carlo-options-simd_b200/build.py --force > /dev/null 2>&1
