#!/bin/bash
# ncu full-set capture of the stage-1+2 kernel for option sets given as arguments: bash tools/ncu_s12.sh <tag> "fuse_planes=0" "fuse_planes=1"
set -u
TAG=$1; shift
i=0
for opts in "$@"; do
  ncu --set full --clock-control none --import-source on -k regex:'k_friedmann' -s 2 -c 1 -f -o gpurun_out/${TAG}_s12_$i python tools/prof_tcgen05.py 7 65536 4 $opts > gpurun_out/${TAG}_s12_$i.log 2>&1
  ncu -i gpurun_out/${TAG}_s12_$i.ncu-rep --page raw --csv > gpurun_out/${TAG}_s12_$i.raw.csv 2>/dev/null
  ncu -i gpurun_out/${TAG}_s12_$i.ncu-rep --page source --csv > gpurun_out/${TAG}_s12_$i.src.csv 2>/dev/null
  i=$((i+1))
done
ls -la gpurun_out/ | head -30
