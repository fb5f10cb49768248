#!/bin/bash
# usage (GPU box): bash tools/gpu_atomic.sh OUTNAME — the shared-atomic scatter, plain and warp-aggregated: rates + ncu
O=gpurun_out/$1; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_splat.py -x -q -m gpu -k "atomic" 2>&1 | tail -2
B="python bench.py --mode atomic --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-extras --no-render-c5"
for a in 0 1; do
  PBRT_B200_ATOMIC_AGG=$a $B > $O/bench_atomic_agg$a.json 2>/dev/null
  echo "agg=$a $(python -c "import json; d=json.loads(open('$O/bench_atomic_agg$a.json').read().strip().splitlines()[-1]); print('%.4g samples/s %.3f ms' % (d['value'], d['ms_per_step']))")" | tee -a $O/summary.txt
  PBRT_B200_ATOMIC_AGG=$a timeout 600 ncu --set full --import-source on --clock-control none -k regex:splat_atomic -c 1 -f -o $O/atomic_agg$a \
     --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum.pct_of_peak_sustained_elapsed,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum \
     $B --steps 1 --warmup 1 > $O/ncu_agg$a.log 2>&1; echo "ncu agg=$a rc=$?" | tee -a $O/summary.txt
done
