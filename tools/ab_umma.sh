#!/bin/bash
# Same-box A/B of two liborx builds (box-to-box clock / power variance is ~10 %, so only
# measurements taken in ONE gpurun call are comparable):  tools/ab_umma.sh liborx.so liborx_roles.so
cd "$(dirname "$0")/.."
for rep in 1 2; do
  for L in "$@"; do
    for cfg in "--dtype bf16 --batch 64 --steps 20" "--dtype bf16 --batch 1024 --steps 10"; do
      export ORX_LIB=$PWD/outline_rag_b200/$L
      timeout 200 python bench.py --rows 6000000 $cfg --no-cpu-baseline --verify 0 2>/dev/null | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); r=d['roofline']; print(os.environ['ORX_LIB'].split('/')[-1], 'B', d['config']['batch'], '|', r['bound'], round(r['frac'],3), round(r['kernel_ms'],3), 'ms | sm_mhz', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
    done
  done
done
