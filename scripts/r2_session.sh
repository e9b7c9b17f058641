#!/bin/bash
# Round 2 GPU session: GPU suite (or a -k subset), bench in both precisions.
#   gpurun --timeout 900 -- 'bash scripts/r2_session.sh <tag> [pytest -k expression]'
TAG=${1:-r2_sx}
KEXPR=${2:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -s -k "$KEXPR" > gpurun_out/${TAG}_pytest.log 2>&1
else
  timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/${TAG}_pytest.log 2>&1
fi
tail -4 gpurun_out/${TAG}_pytest.log
grep -E "render fast chain 320|tcgen05 fp16|full frame" gpurun_out/${TAG}_pytest.log
for PREC in fp16x3 fp16_fp8x; do
timeout 200 python bench.py --precision $PREC > gpurun_out/${TAG}_bench_$PREC.json 2> gpurun_out/${TAG}_bench_$PREC.err
tail -2 gpurun_out/${TAG}_bench_$PREC.err
python - <<PY
import json
f = "${TAG}_bench_$PREC"
try:
    j = json.load(open(f"gpurun_out/{f}.json"))
    r = j["roofline"]
    print(f, "value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "1lane", round(j["config"]["one_frame_at_a_time"]["value"], 1),
          {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in r["hbm_kernels"].items()}, "conv_ms", round(r["kernel_ms_per_step"], 4),
          "frac", round(r["frac"], 3))
    print("   per layer", r["per_layer_ms"])
    print("   parity", {k: v for k, v in (j.get("parity") or {}).items() if not isinstance(v, (dict, str))})
except Exception as e:
    print(f, "failed", e)
PY
done
