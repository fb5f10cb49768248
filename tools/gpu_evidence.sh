#!/bin/bash
# usage (on the GPU box): bash tools/gpu_evidence.sh OUTNAME — the round's evidence set: bench lines, reference arm,
# ncu launch list, full ncu captures of the splat kernel (h = 2 exact / fused, h = 4), per-CTA traces, sanitizers
O=gpurun_out/$1; mkdir -p $O
python bench.py --steps 200 --warmup 5 > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?" | tee -a $O/summary.txt
python bench.py --steps 200 --warmup 5 --mode fma --no-cpu-baseline --no-extras > $O/bench_n1_fma.json 2>/dev/null; echo "bench fma rc=$?" | tee -a $O/summary.txt
python bench.py --steps 50 --warmup 5 --workload c3 --no-cpu-baseline --no-extras --no-render-c5 > $O/bench_c3.json 2>/dev/null; echo "bench c3 rc=$?" | tee -a $O/summary.txt
python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_reference_arm.json 2>/dev/null; echo "reference arm rc=$?" | tee -a $O/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/launches_bench.log 2>&1; echo "launch list rc=$?" | tee -a $O/summary.txt
bash tools/gpu_ncu.sh $1 c2 c2:fma c5 c3
# digests on the box (gpurun brings back at most 64 MiB): the reports themselves are dropped
for k in c2_exact:33177600 c2_fma:33177600 c5_exact:530841600 c3_exact:132710400; do n=${k%%:*}; s=${k##*:}
  (python tools/ncu_summary.py $O/class_$n.ncu-rep $s; echo; echo "== per source line (thread instructions per sample, share of stall samples) =="; python tools/ncu_lines.py $O/class_$n.ncu-rep $s 1.0) > $O/splat_class_$n.txt 2>&1
  rm -f $O/class_$n.ncu-rep
done
PBRT_B200_LIB=$PWD/pbrt_b200/lib/libpbrt_b200_trace.so python tools/cta_trace.py c2 > $O/cta_trace_c2_pdl.txt 2>/dev/null
PBRT_B200_NO_PDL=1 PBRT_B200_LIB=$PWD/pbrt_b200/lib/libpbrt_b200_trace.so python tools/cta_trace.py c2 > $O/cta_trace_c2_isolated.txt 2>/dev/null
PBRT_B200_NO_PDL=1 PBRT_B200_RANK_W=1,1,1,1 PBRT_B200_LIB=$PWD/pbrt_b200/lib/libpbrt_b200_trace.so python tools/cta_trace.py c2 > $O/cta_trace_c2_isolated_equal_segments.txt 2>/dev/null
SEL='not full_size and not back_to_back and not random_configurations'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_splat.py -x -q -m gpu -k "$SEL" > $O/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/summary.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_splat.py -x -q -m gpu -k "c1_64x64 or clipping or exact_pixel_and_half or power_of_two" > $O/racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a $O/summary.txt
tail -n 3 $O/memcheck.log; tail -n 3 $O/racecheck.log
cat $O/summary.txt
