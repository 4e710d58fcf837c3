for cfg in "2 4" "1 4" "4 4" "3 4" "2 3" "2 2"; do
  set -- $cfg
  MCB200_NVCC_FLAGS="-DMCB_QUAD_UNROLL=$1 -DMCB_MIN_CTAS=$2" python monte-carlo-options-simd_b200/build.py --force > /dev/null 2>&1
  for w in c2 c5r c4; do
    python bench.py --workload $w --steps $([ $w = c2 ] && echo 100 || echo 3) --warmup 3 --no-cpu --roof-seconds 0.3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('unroll $1 ctas $2 $w frac %.4f kernel_ms %.4f'%(d['roofline']['frac'], d['roofline']['kernel_ms']))"
  done
done
python monte-carlo-options-simd_b200/build.py --force > /dev/null 2>&1
