# K6 multi-column on 256 MiB inputs: per-row kernels (default) and, with SWEEP=1, the row sweep; one JSON line per case
out=gpurun_out
for wl in cfg2_unquoted cfg3_quoted; do for nc in 1 4 8 16; do for sw in ${SWEEPS:-0 1}; do
  CSVB200_MAT_SWEEP=$sw timeout 120 python tools/profile_mat.py $wl $nc 2>&1 | tail -1
done; done; done | tee $out/mat_ab.jsonl
