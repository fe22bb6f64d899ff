#!/bin/bash
# round-1 closing run (v6d on the big dense levels): tests, launch list of one step, bench lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4400 -c 1700 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/bench_r01_final.json | cut -c1-200
timeout 300 python bench.py --points 3000000 --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | tee gpurun_out/bench_r01_3m.json | cut -c1-200
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | tee gpurun_out/bench_r01_reference.json | cut -c1-200
timeout 300 python tools/batch_scenes.py 16 2>/dev/null | tail -1 | tee gpurun_out/batch_1gpu.json
