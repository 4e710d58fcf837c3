#!/bin/bash
# Sweep of the number of CTA waves a large batch is launched as (MCB200_WAVES) on the GPU box.
for wv in ${WAVES:-1 2 3 4 6 8}; do
  for w in ${WORKLOADS:-c5r c4}; do
    MCB200_WAVES=$wv python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-subs --roof-seconds 0.5 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('waves $wv $w frac %.4f kernel_ms %.4f'%(d['roofline']['frac'], d['roofline']['kernel_ms']))"
  done
done
