#!/bin/bash
# Tile-quantisation experiment: frames batched INTO the conv launches vs independent lanes.
#   gpurun --timeout 900 -- 'bash scripts/exp_batch.sh'
mkdir -p gpurun_out
for cfg in "1 4" "2 2" "2 3" "4 1" "4 2" "8 1" "8 2"; do
  set -- $cfg
  B=$1; L=$2
  timeout 200 python bench.py --batch $B --lanes $L --steps 100 --no-cpu-baseline > gpurun_out/exp_batch_b${B}_l${L}.json 2> gpurun_out/exp_batch_b${B}_l${L}.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/exp_batch_b${B}_l${L}.json"))
    r = j["roofline"]
    print("batch $B lanes $L: value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "conv ms/frame", round(r["kernel_ms_per_step"] / $B, 4),
          "per layer/frame", {k: round(v / $B * 1e3, 1) for k, v in r["per_layer_ms"].items()})
except Exception as e:
    print("batch $B lanes $L failed", e)
PY
done
