#!/bin/bash
# A/B of compile-time variants of the path kernels on the GPU box: rebuilds an alternate library per flag set (the
# production library is untouched) and times the kernels with bench.py.   FLAGSETS="a;b;c" WORKLOADS="c2 c5r c4"
IFS=';'
for flags in ${FLAGSETS:-;-DMCB_OLD_SLOT}; do
  unset IFS
  export MCB200_LIB_NAME=libmcb200_ab.so
  MCB200_NVCC_FLAGS="$flags" python monte-carlo-options-simd_b200/build.py --force > /dev/null 2>&1 || echo "build failed: $flags"
  for w in ${WORKLOADS:-c2 c5r c4}; do
    python bench.py --workload $w --steps $([ $w = c1 -o $w = c2 -o $w = c3 ] && echo 200 || echo 5) --warmup 3 --no-cpu --no-subs --roof-seconds 0.5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('flags [$flags] $w frac %.4f kernel_ms %.4f'%(d['roofline']['frac'], d['roofline']['kernel_ms']))"
  done
  IFS=';'
done
unset IFS
rm -f monte-carlo-options-simd_b200/lib/libmcb200_ab.so
