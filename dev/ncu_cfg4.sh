CMD="python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-others"
$CMD > gpurun_out/r2_plain_cfg4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_cfg4.csv $CMD > gpurun_out/r2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:hamming_select_tc|cand_hist|cand_final|hamming_hist|pack_sign_flat" -s 24 -c 8 -o gpurun_out/prof_r2_cfg4 -f $CMD > gpurun_out/r2_ncu_full.log 2>&1
echo "done rc=$?"
