#!/bin/bash
# Round 2 multi-GPU session (gpurun --gpus N): the N-GPU == 1-GPU equality test executed (not skipped), the default
# bench at N GPUs, and the BASELINE config that belongs to N (4: configs[3] 1280x640 batch 16; 8: configs[4] batch 64).
#   gpurun --gpus N --timeout 900 -- 'bash scripts/r2_multi.sh N <tag>'
N=${1:-2}
TAG=${2:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 400 python -m pytest tests -m gpu -x -q -s -k "multi_gpu" > gpurun_out/${TAG}_n${N}_pytest_multi_gpu.log 2>&1
tail -6 gpurun_out/${TAG}_n${N}_pytest_multi_gpu.log
run() {  # name, extra bench args
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N $2 > gpurun_out/${TAG}_bench_n${N}_$1.json 2> gpurun_out/${TAG}_bench_n${N}_$1.err
  tail -2 gpurun_out/${TAG}_bench_n${N}_$1.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/${TAG}_bench_n${N}_$1.json"))
    print("$1", "N=", j["n_gpus"], "value", round(j["value"], 1), "e2e", round(j["e2e"]["value"], 1), j["config"]["workload"], "|", j["config"]["collective"][:60])
    print("   per rank", j["config"]["timed_regions"]["per_rank_median"])
except Exception as e:
    print("$1 failed", e)
PY
}
run default "--steps 20 --warmup 3"
run default_steps100 "--steps 100"
if [ "$N" = "4" ]; then run config3_1280x640_b16 "--height 640 --width 1280 --batch 4 --steps 30"; fi
if [ "$N" = "8" ]; then run config4_video_b64 "--batch 8 --steps 30"; fi
