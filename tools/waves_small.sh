#!/bin/bash
# Small batches with the wave threshold lowered (tuning knobs MCB200_WAVE_MIN_ROUNDS / MCB200_WAVES): shows why batches with fewer
# than 64 rounds per logical warp stay single-wave (profiles/r2_waves.txt, run 3).
for mr in 64 7 4 2 1; do for wv in 2 3 6; do for w in c2 c3 c1; do
  MCB200_WAVE_MIN_ROUNDS=$mr MCB200_WAVES=$wv python bench.py --workload $w --steps 100 --warmup 3 --no-cpu --no-subs --roof-seconds 0.3 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('min_rounds $mr waves<=$wv $w frac %.4f kernel_ms %.4f value_ms %.4f'%(d['roofline']['frac'], d['roofline']['kernel_ms'], d['ms_per_step']))"
done; done; done
