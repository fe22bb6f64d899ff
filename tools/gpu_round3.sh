#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/batch_scenes.py 16 2>/dev/null | tail -1 | tee gpurun_out/batch_1gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/batch_scenes.py 16 2>/dev/null | tail -1 | tee gpurun_out/batch_2gpu.json
timeout 600 python bench.py --points 3000000 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_3m.json | cut -c1-600
