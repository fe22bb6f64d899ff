#!/bin/bash
# round-1 closing run (decoder wavefront + reversed encoder): tests, bench lines, per-level decode times
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/bench_r01_final.json | cut -c1-160
timeout 300 python bench.py --points 3000000 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_r01_3m.json | cut -c1-160
timeout 300 python tools/batch_scenes.py 16 2>/dev/null | tail -1 | tee gpurun_out/batch_1gpu.json
timeout 300 python tools/wave_times.py > gpurun_out/wave_times.txt 2>&1; tail -12 gpurun_out/wave_times.txt
