#!/bin/bash
# ncu full-set capture of the stage-1+2 kernel of one golden configuration: bash tools/ncu_cfg.sh <tag> <config>
set -u
TAG=$1; CFG=$2
ncu --set full --clock-control none --import-source on -k regex:'k_friedmann' -s 2 -c 1 -f -o gpurun_out/${TAG}_${CFG} python tools/prof_config.py $CFG 65536 4 > gpurun_out/${TAG}_${CFG}.log 2>&1
ncu -i gpurun_out/${TAG}_${CFG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${CFG}.raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_${CFG}.ncu-rep --page source --csv > gpurun_out/${TAG}_${CFG}.src.csv 2>/dev/null
tail -2 gpurun_out/${TAG}_${CFG}.log
