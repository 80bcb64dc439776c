#!/bin/bash
# Everything measured on one GPU for the record: bash tools/final_1gpu.sh <tag>   (under gpurun)
set -u
TAG=$1
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python bench.py --n-sn 1590 --no-alt > gpurun_out/${TAG}_bench_1590.json 2>> gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_reference.json
bash tools/ncu_run.sh ${TAG} > /dev/null 2>&1
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof.raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${TAG}_prof.raw.csv > gpurun_out/${TAG}_ncu_full_summary.txt
python tools/bench_configs.py 65536 > gpurun_out/${TAG}_configs.jsonl 2>/dev/null
python tools/run_nested_config3.py 2>/dev/null | tail -1 > gpurun_out/${TAG}_nested_config3.json
python tools/run_profile_grid_config4.py 100 100 2>/dev/null | tail -1 > gpurun_out/${TAG}_profile_grid_1gpu.json
python tools/run_profile_grid_config4.py 100 100 h0=analytic 2>/dev/null | tail -1 > gpurun_out/${TAG}_profile_grid_1gpu_analytic.json
python tools/small_batch_latency.py > gpurun_out/${TAG}_small_batch.log 2>&1
tail -c 600 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
