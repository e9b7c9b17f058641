#!/bin/bash
# Round 2 GPU session: full GPU suite, bench, A/B of the fused head, launch list.
#   gpurun --timeout 900 -- 'bash scripts/r2_session.sh <tag>'
TAG=${1:-r2_sx}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1
tail -4 gpurun_out/${TAG}_pytest.log
grep "render fast chain" gpurun_out/${TAG}_pytest.log
timeout 200 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
for f in ("${TAG}_bench",):
    try:
        j = json.load(open(f"gpurun_out/{f}.json"))
        r = j["roofline"]
        print(f, "value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "1lane", round(j["config"]["one_frame_at_a_time"]["value"], 1),
              {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in r["hbm_kernels"].items()}, "conv_ms", round(r["kernel_ms_per_step"], 4),
              "head", r["per_layer_ms"].get("color_pred"), "launches", j["gpu_launches"])
    except Exception as e:
        print(f, "failed", e)
PY
