#!/bin/bash
# round-1 final measurements (run under gpurun): ncu of the two conv kernels on the 442 133-row level, launch list of one bench step, bench lines
mkdir -p gpurun_out
export AB_MIN_ROWS=400000 AB_MAX_ROWS=500000
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spconv_fwd_v6 -s 2 -c 1 -f -o gpurun_out/r01_v6_final python tools/conv_ab.py 1000000 42:64 > gpurun_out/ncu_v6.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spconv_tc -s 2 -c 1 -f -o gpurun_out/r01_tc_final python tools/conv_ab.py 1000000 100:1024 > gpurun_out/ncu_tc.log 2>&1
unset AB_MIN_ROWS AB_MAX_ROWS
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4700 -c 1600 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 | tee gpurun_out/bench_r01_final.json | cut -c1-200
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | tee gpurun_out/bench_r01_reference.json | cut -c1-200
