#!/bin/bash
# usage: tools/sass_probe.sh H [extra -D...]: compile splat_class.cu with one instantiation <H,128,exact> and print
# the instruction mix of its SASS (static) — MOV counts in the hot loops show up here before any GPU time is spent
H=$1; shift
mkdir -p /tmp/probe
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xptxas=-v -DPBRT_CLASS_PROBE=$H "$@" \
  -c pbrt_b200/csrc/splat_class.cu -o /tmp/probe/p.o 2>&1 | grep -A1 "splat_class_kernel" | grep -E "registers|spill"
cuobjdump -sass /tmp/probe/p.o | grep -E "^\s+/\*[0-9a-f]{4,6}\*/" | sed -E 's#^\s+/\*([0-9a-f]+)\*/\s+#\1 #; s#\s*/\* 0x[0-9a-f]+ \*/##' > /tmp/probe/p.sass
wc -l < /tmp/probe/p.sass
