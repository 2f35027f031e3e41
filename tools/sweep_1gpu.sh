#!/bin/bash
# One-GPU sweep over the BASELINE.json configs (writes JSON lines): tools/sweep_1gpu.sh out.jsonl
cd "$(dirname "$0")/.."
out=${1:-gpurun_out/sweep.jsonl}; : > "$out"
for cfg in "--rows 1000000 --dtype fp32 --batch 1 --steps 200" "--rows 1000000 --dtype fp32 --batch 64 --steps 50" \
           "--rows 10000000 --dtype fp32 --batch 1 --steps 100" "--rows 10000000 --dtype fp32 --batch 8 --steps 30" \
           "--rows 10000000 --dtype fp32 --batch 64 --steps 30" "--rows 10000000 --dtype fp32 --batch 256 --steps 20" \
           "--rows 10000000 --dtype fp32 --batch 1024 --steps 8" "--rows 10000000 --dtype bf16 --batch 1 --steps 100" \
           "--rows 10000000 --dtype bf16 --batch 64 --steps 30" "--rows 10000000 --dtype bf16 --batch 256 --steps 20" \
           "--rows 10000000 --dtype bf16 --batch 1024 --steps 10 --recall-queries 128"; do
  timeout 300 python bench.py $cfg --no-cpu-baseline 2>/dev/null | tail -1 >> "$out"
done
python - "$out" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    d=json.loads(l); r=d["roofline"]
    print(d["config"]["workload"][22:], "| ms/step", round(d["ms_per_step"],3), "| qps", round(d["value"],1), "| p50", round(d["p50_ms"],3), "|", r["bound"], round(r["frac"],3), r["kernel"], round(r["kernel_ms"],3), "| e2e", round(d["e2e"]["value"],1), d["fallbacks"], d["verify"].get("sampled_rescoring_bit_exact"), d.get("recall"))
PY
