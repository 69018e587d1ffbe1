#!/bin/bash
# round-2 ncu evidence: launch list of a bench run + full captures of the hot kernels of one late step
mkdir -p gpurun_out
W=${1:-c2}
CMD="python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-config3 --no-gpu-baseline"
$CMD > gpurun_out/ncu_plain_$W.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$W.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_$W.csv $CMD > gpurun_out/ncu_list_$W.log 2>&1
echo "ncu list exit $?"
$CMD > gpurun_out/ncu_plain2_$W.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"tc_kernel|select_kernel|rerank_kernel|output_kernel|baseprep|tau_merge|qprep" -s ${SKIP:-65} -c ${COUNT:-13} -f -o gpurun_out/r02_prof_$W $CMD > gpurun_out/ncu_full_$W.log 2>&1
echo "ncu full exit $?"
tail -2 gpurun_out/ncu_full_$W.log | cut -c1-200
ncu -i gpurun_out/r02_prof_$W.ncu-rep --page raw --csv > gpurun_out/r02_ncu_raw_$W.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r02_launches_$W.csv')))
for i,r in enumerate(rows):
    if r and r[0]=='ID': hdr=r; start=i+1; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
recs=[(r[ki][:70], float(r[vi].replace(',',''))) for r in rows[start:] if len(r)>vi]
for k,v in recs[-16:]: print(f'{v/1e6:9.3f} ms  {k}')
PY
