# usage (under gpurun): bash dev/ncu_cfg4.sh [tag]   -- plain run, launch list, one whole step under --set full (cfg4),
# and the two select launches (sample + full gallery) of one cfg5 step
TAG="${1:-r2}"
CMD="python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-others"
$CMD > gpurun_out/${TAG}_plain_cfg4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_cfg4.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:hamming_select_tc|cand_hist|cand_final|cand_rank|hamming_hist|pack_sign_flat|expand_i8" -s 30 -c 10 -o gpurun_out/prof_${TAG}_cfg4 -f $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "cfg4 done rc=$?"
CMD5="python bench.py --workload cfg5 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-others"
$CMD5 > gpurun_out/${TAG}_plain_cfg5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:hamming_select_tc" -s 6 -c 2 -o gpurun_out/prof_${TAG}_cfg5 -f $CMD5 > gpurun_out/${TAG}_ncu_full5.log 2>&1
echo "cfg5 done rc=$?"
