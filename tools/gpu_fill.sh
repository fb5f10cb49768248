#!/bin/bash
O=gpurun_out/r2fill; mkdir -p $O
for v in 0 1 2; do for c in 16 32 64; do PBRT_B200_FILL_VARIANT=$v PBRT_B200_FILL_CTAS=$c python tools/fill_bench.py 2>/dev/null | tee -a $O/fill.jsonl; done; done
