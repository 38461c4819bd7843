#!/bin/bash
# Runs ON the GPU box (under gpurun): ncu evidence for every kernel, reduced to text so it fits the 64 MiB return path.
#   gpurun_out/${tag}_ncu_all_kernels.md     one row per launch of every kernel (tools/ncu_table.py)
#   gpurun_out/prof_r1_v6_cfg{2,3}.ncu-rep   full capture (with source) of the hot kernel on the 1 GiB configs
#   gpurun_out/launches_${tag}.csv            launch list of `bench.py` (gpu__time_duration)
set -u
out=gpurun_out
tag=${1:-r02}
python tools/profile_all.py > $out/prof_all_plain.log 2>&1 || { tail -5 $out/prof_all_plain.log; exit 1; }
ncu --set full --clock-control none -f -o /tmp/prof_all python tools/profile_all.py > $out/prof_all_ncu.log 2>&1
python tools/ncu_table.py /tmp/prof_all.ncu-rep > $out/${tag}_ncu_all_kernels.md 2> $out/ncu_table.err
wc -l $out/${tag}_ncu_all_kernels.md
for cfg in cfg2_unquoted cfg3_quoted; do
  short=${cfg%%_*}
  python tools/profile_run.py $cfg 4 > $out/prof_${short}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:index_build_tma_kernel -s 2 -c 1 -f -o $out/prof_${tag}_$short \
      python tools/profile_run.py $cfg 4 > $out/prof_${short}_ncu.log 2>&1
  tail -1 $out/prof_${short}_plain.log
done
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_${tag}.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/ncu_launch.log 2>&1
grep -c . $out/launches_${tag}.csv
du -sh $out
