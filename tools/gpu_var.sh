#!/bin/bash
# usage: tools/gpu_var.sh OUTNAME variant[:workload[:mode]] ...   (variant "base" = the product library)
# one short device-resident bench line per variant library (pbrt_b200/lib/libpbrt_b200_<variant>.so)
O=gpurun_out/$1; shift; mkdir -p $O
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-extras --no-render-c5"
for spec in "$@"; do
  IFS=: read v wl mode <<< "$spec"; wl=${wl:-c2}; mode=${mode:-exact}
  lib=$PWD/pbrt_b200/lib/libpbrt_b200_$v.so; [ "$v" = base ] && lib=$PWD/pbrt_b200/lib/libpbrt_b200.so
  steps=100; [ "$wl" = c5 ] && steps=20
  name=${v}_${wl}_${mode}
  PBRT_B200_LIB=$lib $B --workload $wl --mode $mode --steps $steps > $O/$name.json 2> $O/$name.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$O/$name.json').read().strip().splitlines()[-1]); print('%.4g %.4f ms frac %.4f' % (d['value'], d['ms_per_step'], d['roofline']['frac']))
except Exception as e: print('parse-fail', e)
PY
)" | tee -a $O/summary.txt
done
