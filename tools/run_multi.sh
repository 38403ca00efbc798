#!/bin/bash
# Multi-GPU validation + measurements on ONE box: tools/run_multi.sh N   (writes gpurun_out/*_${N}gpu_r02.*)
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
mkdir -p gpurun_out
timeout 400 $TR tools/dist_check.py > gpurun_out/dist_check_${N}gpu_r02.log 2>&1; echo "dist_check rc=$?"; grep -c OK gpurun_out/dist_check_${N}gpu_r02.log; tail -1 gpurun_out/dist_check_${N}gpu_r02.log
timeout 500 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu 2> gpurun_out/bench_${N}gpu_r02.err | tail -1 > gpurun_out/bench_${N}gpu_r02.json; echo "bench rc=$?"
timeout 400 $TR tools/bench_cfg4.py 2> gpurun_out/cfg4_${N}gpu_r02.err | tail -1 > gpurun_out/cfg4_${N}gpu_r02.json; echo "cfg4 rc=$?"
if [ "$N" = "8" ]; then
  timeout 600 $TR tools/bench_cfg5.py 2> gpurun_out/cfg5_${N}gpu_r02.err | tail -1 > gpurun_out/cfg5_${N}gpu_r02.json; echo "cfg5 rc=$?"
fi
