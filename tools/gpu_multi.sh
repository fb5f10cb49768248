#!/bin/bash
# usage (on a multi-GPU box): bash tools/gpu_multi.sh OUTNAME — multi-GPU tests and the bench line at N = all GPUs
O=gpurun_out/$1; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $O/pytest_multi.log 2>&1; echo "pytest_multi rc=$?" | tee -a $O/summary.txt
tail -3 $O/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench_n$N rc=$?" | tee -a $O/summary.txt
tail -3 $O/bench_n$N.err
python - <<PY
import json
d=json.loads(open('$O/bench_n$N.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e'], d['assemble'])
print(json.dumps(d['render_c5'], indent=1))
PY
