O=gpurun_out/r2u; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_splat.py -x -q -m gpu 2>&1 | tail -2
PBRT_B200_LIB=$PWD/pbrt_b200/lib/libpbrt_b200_trace.so python tools/cta_trace.py c2 2>/dev/null | tee $O/trace_c2.txt
B="python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-extras --no-render-c5"
for w in "1,1,1,1" "1,0.95,0.90,0.865" "1,0.93,0.87,0.82" "1,0.96,0.92,0.89" "1,0.9,0.82,0.76"; do
  for c in 1.2; do
    echo "W=$w C=$c $(PBRT_B200_RANK_W=$w PBRT_B200_HALO_C=$c $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g %.4f' % (d['value'], d['ms_per_step']))")" | tee -a $O/summary.txt
  done
done
for wl in c3 c5; do for w in "1,1,1,1" "1,0.95,0.90,0.865" "1,0.93,0.87,0.82"; do
  echo "$wl W=$w $(PBRT_B200_RANK_W=$w $B --workload $wl --steps 20 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.4g %.4f' % (d['value'], d['ms_per_step']))")" | tee -a $O/summary.txt
done; done
