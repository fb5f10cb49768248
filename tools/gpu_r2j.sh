#!/bin/bash
O=gpurun_out/r2j; mkdir -p $O
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest_gpu rc=$?" | tee -a $O/summary.txt
tail -5 $O/pytest_gpu.log
python - <<PY
import json
d=json.loads(open('$O/bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e'])
print(json.dumps(d['render_c5'], indent=1))
PY
