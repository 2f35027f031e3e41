#!/bin/bash
# Same-box A/B of the single-query scan for two liborx builds: tools/ab_gemv.sh liborx_base.so liborx.so
cd "$(dirname "$0")/.."
for rep in 1 2; do
  for L in "$@"; do
    for cfg in "--rows 6000000 --dtype bf16 --batch 1 --steps 100" "--rows 3000000 --dtype fp32 --batch 1 --steps 100"; do
      export ORX_LIB=$PWD/outline_rag_b200/$L
      timeout 200 python bench.py $cfg --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); r=d['roofline']; print(os.environ['ORX_LIB'].split('/')[-1], d['config']['workload'][22:], '|', r['bound'], round(r['frac'],4), round(r['kernel_ms'],4), 'ms step', round(d['ms_per_step'],4), d['verify'])"
    done
  done
done
