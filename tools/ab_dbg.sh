#!/bin/bash
# Same-box sweep of ORX_UMMA_DEBUG values: tools/ab_dbg.sh "<bench args>" v1 v2 ...
cd "$(dirname "$0")/.."
args="$1"; shift
for rep in 1 2; do
  for D in "$@"; do
    export ORX_UMMA_DEBUG=$D
    timeout 200 python bench.py $args --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); r=d['roofline']; print('dbg', os.environ['ORX_UMMA_DEBUG'], '|', r['bound'], round(r['frac'],3), round(r['kernel_ms'],3), 'ms', d['fallbacks'], d['verify'], d['clocks']['sm_mhz'])"
  done
done
