#!/bin/bash
# full-set ncu capture of the stage-3 GEMM and stage-1/2 kernel (one launch each) after a plain run
TAG=${1:-r01b}; shift
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline $@"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:'k_(friedmann|chi2)' -s 6 -c 2 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
tail -3 gpurun_out/${TAG}_ncu2.log
