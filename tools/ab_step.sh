#!/bin/bash
# Same-box A/B of whole search steps for two liborx builds: tools/ab_step.sh liborx_old.so liborx.so
cd "$(dirname "$0")/.."
for rep in 1 2; do
  for L in "$@"; do
    for cfg in "--rows 2000000 --dtype bf16 --batch 1024 --steps 10" "--rows 2000000 --dtype fp32 --batch 256 --steps 10" "--rows 1000000 --dtype fp32 --batch 1 --steps 200"; do
      export ORX_LIB=$PWD/outline_rag_b200/$L
      timeout 200 python bench.py $cfg --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); r=d['roofline']; print(os.environ['ORX_LIB'].split('/')[-1], d['config']['workload'][22:], '| step', round(d['ms_per_step'],4), 'scan', round(r['kernel_ms'],4), d['fallbacks'], d['verify'])"
    done
  done
done
