#!/bin/bash
# times the contraction for every library variant under variants/ (experiments with -DOZ_PHASES / -DOZ_HEAD) and dumps traces
for lib in cosmology_model_fit_b200/libcosmolike_b200.so variants/*.so; do
  for S in 7 6; do
    echo "== $lib S=$S"
    COSMOLIKE_LIB=$PWD/$lib timeout 120 python tools/prof_tcgen05.py $S 65536 6 2>&1 | tail -1 | sed 's/.*split/split/'
    [ "$lib" != variants/lib_base.so ] && COSMOLIKE_TRACE=gpurun_out/trace_$(basename $lib .so)_s$S.txt COSMOLIKE_LIB=$PWD/$lib timeout 120 python tools/prof_tcgen05.py $S 65536 2 dbg=12 2>&1 | grep "oz prof" | tail -1
  done
done
