#!/bin/bash
# Same-box A/B of liborx builds (box-to-box clock / power variance is ~10 %, so only measurements taken in ONE
# gpurun call are comparable).  Variant builds: make -C outline_rag_b200/csrc variant NAME=x DEFS=-D...
#   tools/ab.sh "<bench args>" liborx.so liborx_x.so ...      e.g.  tools/ab.sh "--rows 6000000 --dtype bf16 --batch 1024 --steps 10"
# Timing experiments of the tcgen05 scan exist only in -DORX_DEBUG_VARIANTS builds (ORX_UMMA_DEBUG=<bits> then applies).
cd "$(dirname "$0")/.."
ARGS="$1"; shift
for rep in 1 2; do
  for L in "$@"; do
    timeout 300 python bench.py $ARGS --lib outline_rag_b200/$L --no-cpu-baseline --verify 0 --configs none 2>/dev/null | tail -1 | L=$L python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); r=d['roofline']
print(os.environ['L'], d['config']['workload'][22:], '|', r['bound'], 'frac', round(r['frac'],3), 'kernel', round(r['kernel_ms'],4), 'ms | step', round(d['ms_per_step'],4), 'ms | p50', round(d['latency']['p50_ms'],4), '| sm_mhz', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
  done
done
