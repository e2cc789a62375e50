set -x
CMD="python bench.py --workload lbl --steps 2 --warmup 3 --no-cpu-baseline --no-parity"
$CMD > gpurun_out/r2_plain_lbl.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:rcm_lbl_rt_kernel -s 3 -c 1 -f -o gpurun_out/r2_lbl $CMD > gpurun_out/r2_ncu_b.log 2>&1
ls -la gpurun_out/r2_lbl.ncu-rep
