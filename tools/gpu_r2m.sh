#!/bin/bash
O=gpurun_out/r2m; mkdir -p $O
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-extras --no-render-c5"
run() { name=$1; shift; ( "$@" ) > $O/$name.json 2> $O/$name.err; echo "$name rc=$? $(python - <<PY
import json,sys
try:
    d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['frac'])
except Exception as e: print('parse-fail', e)
PY
)" | tee -a $O/summary.txt; }
run c2_class $B
run c2_class_fma $B --mode fma
run c3_class $B --workload c3
run c5_class $B --workload c5 --steps 20
timeout 600 ncu --set full --import-source on --clock-control none -k regex:splat_class_kernel -c 1 -o $O/class_c2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --no-render-c5 > $O/ncu_c2.log 2>&1; echo "ncu c2 rc=$?" | tee -a $O/summary.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:splat_class_kernel -c 1 -o $O/class_c5 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --no-render-c5 > $O/ncu_c5.log 2>&1; echo "ncu c5 rc=$?" | tee -a $O/summary.txt
cat $O/summary.txt
