#!/bin/bash
# BASELINE.json configs[2]: 640x320 ERP, 64-sphere MSI, batch 8 on one B200 (deep-layer composite stress)
mkdir -p gpurun_out
timeout 500 python bench.py --planes 64 --batch 8 --steps 20 --lanes 2 > gpurun_out/r2_bench_config2_p64_b8.json 2> gpurun_out/r2_bench_config2_p64_b8.err
tail -3 gpurun_out/r2_bench_config2_p64_b8.err
python - <<PY
import json
j = json.load(open("gpurun_out/r2_bench_config2_p64_b8.json")); r = j["roofline"]
print("config2", j["config"]["workload"], "value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), "conv_ms", round(r["kernel_ms_per_step"], 3), "frac", round(r["frac"], 3),
      {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in r["hbm_kernels"].items()}, "cpu", j["cpu_baseline"] and round(j["cpu_baseline"]["value"], 3))
PY
