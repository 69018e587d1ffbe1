#!/bin/bash
# launch list + full capture of selected kernels (regex in $1) for workload $2
mkdir -p gpurun_out
K=${1:-rerank}
W=${2:-c2}
CMD="python bench.py --workload $W --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$W.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on -k regex:"$K" -s 6 -c 3 -f -o gpurun_out/prof_sel_$W $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/launches_$W.csv')))
for i,r in enumerate(rows):
    if r and r[0]=='ID': hdr=r; start=i+1; break
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
recs=[(r[ki][:60], float(r[vi].replace(',',''))) for r in rows[start:] if len(r)>vi]
for k,v in recs[-12:]: print(f'{v/1e6:9.3f} ms  {k}')
PY
