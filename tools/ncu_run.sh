#!/bin/bash
# ncu evidence for the round: launch list (per-launch durations) + one full-set capture of each hot kernel.
# Usage (under gpurun): bash tools/ncu_run.sh <tag>
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-alt"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_(friedmann|chi2|finalize|oz)' -c 30 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_(friedmann|chi2_ozaki)' -s 6 -c 2 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/
