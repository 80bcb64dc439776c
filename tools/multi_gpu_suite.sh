#!/bin/bash
# Everything measured at N GPUs for the record: bash tools/multi_gpu_suite.sh <tag> <N>   (under gpurun --gpus N)
set -u
TAG=$1; N=$2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$TR --master-port 29511 tools/multi_gpu_check.py 2>/dev/null | tail -1 > gpurun_out/${TAG}_check_${N}gpu.json
$TR --master-port 29512 bench.py --gpus $N --no-alt > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_${N}gpu.err
$TR --master-port 29513 bench.py --gpus $N --no-alt --scaling strong > gpurun_out/${TAG}_bench_${N}gpu_strong.json 2>> gpurun_out/${TAG}_${N}gpu.err
$TR --master-port 29514 bench.py --gpus $N --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_${N}gpu_reference.json
$TR --master-port 29515 tools/run_profile_grid_config4.py 100 100 2>/dev/null | tail -1 > gpurun_out/${TAG}_profile_grid_${N}gpu.json
python - <<PY
import json
for f in ("bench_${N}gpu", "bench_${N}gpu_strong"):
    try:
        d = json.loads([l for l in open("gpurun_out/${TAG}_%s.json" % f).read().splitlines() if l.startswith("{")][-1])
        print(f, d["scaling"], "value %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"],
              "e2e_all %.4g" % d.get("e2e_all_ranks", {}).get("value", 0), "batch/gpu", d["config"]["batch_per_gpu"], d["clocks"]["reasons"])
    except Exception as e:
        print(f, "FAILED", e)
for f in ("check_${N}gpu", "bench_${N}gpu_reference", "profile_grid_${N}gpu"):
    try:
        print(open("gpurun_out/${TAG}_%s.json" % f).read()[:400])
    except Exception as e:
        print(f, "FAILED", e)
PY
