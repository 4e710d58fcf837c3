set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_box.txt; nproc >> gpurun_out/r2_box.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/r2_box.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu_1.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_gpu_1.log
tail -5 gpurun_out/r2_pytest_gpu_1.log
timeout 300 python tools/microbench.py > gpurun_out/r2_microbench_tri.json 2> gpurun_out/r2_microbench_tri.err
for w in c1 c2 c3 c5r c4; do
  timeout 300 python bench.py --workload $w --steps $([ $w = c1 -o $w = c2 -o $w = c3 ] && echo 100 || echo 3) --warmup 3 --no-cpu --roof-seconds 0.5 2>gpurun_out/r2_bench1_$w.err | tail -1 > gpurun_out/r2_bench1_$w.json
  python -c "import json; d=json.load(open('gpurun_out/r2_bench1_$w.json')); print('$w frac %.4f kernel_ms %.4f e2e %.4g'%(d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value']))"
done
timeout 600 python tools/kernel_timeline.py c1 c2 c3 > gpurun_out/r2_timeline.log 2>&1
tail -3 gpurun_out/r2_timeline.log
timeout 1500 tools/run_sanitizer.sh price crossing greeks moments
