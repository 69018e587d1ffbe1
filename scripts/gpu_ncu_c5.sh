CMD="python bench.py --workload c5shard --steps 1 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_kernel" -s 6 -c 2 -o gpurun_out/prof_c5 $CMD > gpurun_out/ncu_full_c5.log 2>&1
echo "ncu full exit $?"
