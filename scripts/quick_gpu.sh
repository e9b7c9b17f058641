#!/bin/bash
# quick GPU check: a pytest -k subset and a short bench.   gpurun --timeout 900 -- 'bash scripts/quick_gpu.sh <tag> "<k expr>" [bench args]'
TAG=${1:-q}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -s -k "$2" > gpurun_out/${TAG}_pytest.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -15
timeout 200 python bench.py --steps 50 --no-cpu-baseline $3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
try:
    j = json.load(open("gpurun_out/${TAG}_bench.json")); r = j["roofline"]
    print("value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "1lane", round(j["config"]["one_frame_at_a_time"]["value"], 1),
          "conv_ms", round(r["kernel_ms_per_step"], 4), "frac", round(r["frac"], 3), "clk", j["clocks"]["sm_mhz"])
    print("  per layer us", {k: round(v * 1e3, 1) for k, v in r["per_layer_ms"].items()})
    print("  energy", j.get("energy"))
    print("  parity", {k: v for k, v in (j.get("parity") or {}).items() if not isinstance(v, (dict, str))})
except Exception as e:
    print("bench failed", e)
PY
