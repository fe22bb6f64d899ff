#!/bin/bash
# round-1 closing run: GPU test-suite, config[2] sweep, bench lines (N=1 here; N=2 separately under gpurun --gpus 2)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/sweep.py 1,2,4,8,16,32 2>/dev/null | tee gpurun_out/sweep_r01_final.jsonl | cut -c1-120
timeout 600 python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/bench_r01_final.json | cut -c1-200
timeout 300 python bench.py --points 3000000 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_r01_3m.json | cut -c1-200
