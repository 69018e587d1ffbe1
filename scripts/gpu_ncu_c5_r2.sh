#!/bin/bash
# --set full capture of the two tc_kernel launches (sample, filter) of one late step of the config-5 shard, final round-2 build
mkdir -p gpurun_out
CMD="python bench.py --workload c5shard --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-config3 --no-gpu-baseline"
$CMD > gpurun_out/ncu_plain_c5.log 2>&1 || { echo "plain run failed"; tail -3 gpurun_out/ncu_plain_c5.log; exit 1; }
# a step launches tc_kernel 4 times (sample, filter, and the two of the second chance): skip 3 steps = 12 launches
ncu --set full --clock-control none --import-source on -k regex:"tc_kernel" -s 12 -c 2 -f -o gpurun_out/r02_prof_c5tc $CMD > gpurun_out/ncu_full_c5tc.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full_c5tc.log | cut -c1-200
ncu -i gpurun_out/r02_prof_c5tc.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw_c5_tc_kernels.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r02_ncu_raw_c5_tc_kernels.csv')))
hdr=rows[0]
def col(n): return hdr.index(n)
for r in rows[2:]:
    print(r[col('Kernel Name')][:60], r[col('Grid Size')], 'ms', r[col('gpu__time_duration.sum')],
          'tensor%', r[col('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed')],
          'dram rd/wr', r[col('dram__bytes_read.sum')], r[col('dram__bytes_write.sum')])
PY
