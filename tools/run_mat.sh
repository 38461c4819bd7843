# K6 multi-column: per-row kernels (CSVB200_MAT_SWEEP=0) against the row sweep, then one ncu capture of the sweep
out=gpurun_out
for wl in cfg2_unquoted cfg3_quoted; do for nc in 4 8 16; do for sw in 0 1; do
  CSVB200_MAT_SWEEP=$sw timeout 120 python tools/profile_mat.py $wl $nc 2>&1 | tail -1
done; done; done | tee $out/mat_ab.jsonl
CSVB200_MAT_SWEEP=1 ncu --set full --clock-control none --import-source on -k regex:sweep -s 4 -c 2 -f -o $out/prof_mat_cfg2 python tools/profile_mat.py cfg2_unquoted 8 > $out/prof_mat_ncu.log 2>&1
ls -la $out/prof_mat_cfg2.ncu-rep
