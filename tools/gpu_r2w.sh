O=gpurun_out/r2w; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_splat.py -x -q -m gpu 2>&1 | tail -2
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-extras --no-render-c5"
run() { echo "$1 $( ( shift; "$@" ) 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g %.4f' % (d['value'], d['ms_per_step']))")" | tee -a $O/summary.txt; }
run nopdl env PBRT_B200_NO_PDL=1 $B
run pdl_waitfirst env PBRT_B200_PDL_WAIT_FIRST=1 $B
run pdl $B
for w in "1,1,1,1" "1,0.97,0.94,0.91" "1,0.95,0.90,0.85" "1,0.93,0.87,0.82"; do run "pdl_W=$w" env PBRT_B200_RANK_W=$w $B; done
run pdl_c3 $B --workload c3
run pdl_c3_uniform env PBRT_B200_RANK_W=1,1,1,1 $B --workload c3
run pdl_c5 $B --workload c5 --steps 20
run pdl_fma $B --mode fma
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['render_c5']['ms'], d['render_c5']['splat_ms'])"
