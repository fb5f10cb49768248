#!/bin/bash
# usage: tools/gpu_ncu.sh OUTNAME [workload:mode ...] — one `ncu --set full` capture of the splat kernel per entry
O=gpurun_out/$1; shift; mkdir -p $O
for spec in "$@"; do
  IFS=: read wl mode <<< "$spec"; mode=${mode:-exact}
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:splat_class_kernel -c 1 -o $O/class_${wl}_${mode} -f \
    python bench.py --workload $wl --mode $mode --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-extras --no-render-c5 > $O/ncu_${wl}_${mode}.log 2>&1
  echo "ncu $wl $mode rc=$?" | tee -a $O/summary.txt
done
