#!/bin/bash
# Round-2 evidence capture on one B200 (run through gpurun from the repo root). Writes into gpurun_out/:
#   r2_launches_B<b>.csv        ncu launch list (time + DRAM bytes) of one warm-up + one timed Stage-I step
#   r2_full_<kernel>.ncu-rep    one `--set full` capture per kernel of interest (second matching launch)
# Nothing printed by a run under ncu is a bench value.
B=${1:-4096}
OUT=gpurun_out
NCU="ncu --clock-control none"
[ -n "$SKIP_LAUNCH_LIST" ] || $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file $OUT/r2_launches_B$B.csv \
    python bench.py --steps 1 --warmup 1 --quick --batch $B > $OUT/r2_ncu_launches_B$B.log 2>&1
# regexes over the DEMANGLED names, which spell template arguments as `<(int)256, (int)64, ...>` ('.' stands for <, ( and ))
KERNELS=${KERNELS:-"igemm_persistent_kernel..int.256;igemm_persistent_kernel..int.128, .int.32;igemm_persistent_kernel..int.128, .int.64;wgrad_kernel..int.128, .int.64, .int.2;wgrad_kernel..int.32, .int.32;hconv_kernel..int.32;cto3_kernel;hwgrad_kernel;bn_bwd_cp_kernel..int.1, __nv_bf;bn_bwd_cp_kernel..int.2, __nv_bf;bn_apply_kernel;mt_rmsprop_kernel;rowsqdiff_fwd_kernel"}
IFS=';' read -ra KS <<< "$KERNELS"
for K in "${KS[@]}"; do
  F=$(echo "$K" | tr -c 'A-Za-z0-9_' '_')
  $NCU --set full --kernel-name-base demangled -k regex:"$(echo "$K" | sed 's/[<>,]/./g')" -s 1 -c 1 -f -o $OUT/r2_full_$F \
      python bench.py --steps 1 --warmup 0 --quick --batch $B > $OUT/r2_ncu_full_$F.log 2>&1
  # the report embeds the whole cubin (16 MB): keep its raw metric page (one CSV row per captured launch) and drop the report
  ncu -i $OUT/r2_full_$F.ncu-rep --page raw --csv > $OUT/r2_full_$F.csv 2>/dev/null
  rm -f $OUT/r2_full_$F.ncu-rep
done
ls -la $OUT/r2_full_*.csv
