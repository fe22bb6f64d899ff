#!/bin/bash
# 2-GPU check of the torchrun path + reference arm
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.log 2>&1; echo "rc=$?" >> gpurun_out/bench2.log
tail -4 gpurun_out/bench2.log
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; tail -2 gpurun_out/bench_ref.log
