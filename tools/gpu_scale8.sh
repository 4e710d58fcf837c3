#!/bin/bash
# 8-GPU session: strong-scaling curve of bench.py (one process per GPU), the in-library multi-device context, the
# two-GPU tests and the fused-vs-NCCL trials-minor comparison.   gpurun --gpus 8 -- tools/gpu_scale8.sh
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 --no-subs --no-cpu > gpurun_out/r2_scale_n1.json 2> gpurun_out/r2_scale_n1.err; echo "N=1 rc $?"
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_scale_n$N.json 2> gpurun_out/r2_scale_n$N.err; echo "N=$N rc $?"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2_scale_ref_n8.json 2> gpurun_out/r2_scale_ref_n8.err; echo "ref N=8 rc $?"
timeout 600 python tools/multi_ctx_bench.py > gpurun_out/r2_multi_ctx_bench.log 2>&1; echo "multi ctx rc $?"
timeout 600 python -m pytest tests -m gpu -q -k multigpu > gpurun_out/r2_pytest_multigpu_8gpu.log 2>&1; echo "pytest rc $?"
for T in 8e6 8e5; do
  XGPU_TRIALS=$T python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29630 tools/xgpu_bench.py > gpurun_out/r2_xgpu_n8_$T.json 2> gpurun_out/r2_xgpu_n8_$T.err; echo "xgpu $T rc $?"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29640 bench.py --gpus 8 --steps 20 --warmup 5 --workload c4 > gpurun_out/r2_scale_c4_n8.json 2> gpurun_out/r2_scale_c4_n8.err; echo "c4 N=8 rc $?"
