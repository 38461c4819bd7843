#!/bin/bash
# ncu --set full (with source) of the hot kernel under several CSVB200_TUNE values on one workload.
# usage (GPU box): bash tools/profile_variants.sh cfg3_quoted "8192 9216" tag
wl=$1; tunes=$2; tag=${3:-pv}
out=gpurun_out
for t in $tunes; do
  CSVB200_TUNE=$t python tools/profile_run.py $wl 4 > $out/${tag}_t${t}_plain.log 2>&1 &&
  CSVB200_TUNE=$t ncu --set full --clock-control none --import-source on -k regex:index_build_tma_kernel -s 3 -c 1 -f \
      -o $out/${tag}_${wl%%_*}_t$t python tools/profile_run.py $wl 4 > $out/${tag}_t${t}_ncu.log 2>&1
  tail -1 $out/${tag}_t${t}_plain.log
done
ls -la $out/*.ncu-rep | tail -5
