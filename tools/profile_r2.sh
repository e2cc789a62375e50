set -x
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-strong"
$CMD > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rcm_split_rt_kernel -s 3 -c 1 -f -o gpurun_out/r2_step $CMD > gpurun_out/r2_ncu_s.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rcm_lbl_rt_kernel -s 3 -c 1 -f -o gpurun_out/r2_lbl $CMD > gpurun_out/r2_ncu_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rcm_split_col_kernel -s 4 -c 1 -f -o gpurun_out/r2_col $CMD > gpurun_out/r2_ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/r2_ncu_s.log
