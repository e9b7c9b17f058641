#!/bin/bash
# What the round-1 GPU budget did not reach for the CTA-pair conv kernel (DESIGN.md sect. 8).  One GPU:
#   gpurun --timeout 900 -- 'bash scripts/next_gpu_session.sh'
# then (separately, N GPUs):  gpurun --gpus 8 --timeout 600 -- 'python -m torch.distributed.run --nnodes=1 \
#   --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/r2_bench_n8.json'
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 280 compute-sanitizer --tool $tool python scripts/sanitize_small.py > gpurun_out/r2_compute_sanitizer_$tool.log 2>&1
  tail -2 gpurun_out/r2_compute_sanitizer_$tool.log
done
# pipeline-event clocks of CTA 0 (leader of pair 0) in conv2_1: per-slot waits, unit / chunk boundaries
MSI_TC_TRACE=conv2_1 timeout 120 python scripts/trace_conv.py > gpurun_out/r2_trace_conv2_1_pair.log 2>&1
tail -5 gpurun_out/r2_trace_conv2_1_pair.log
# all 18 conv launches of one forward (the round-1 capture stopped after 13: ncu needs ~12 s per launch here)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:conv_ -s 18 -c 18 -f -o gpurun_out/r2_conv \
    python scripts/one_forward.py > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log
timeout 120 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
timeout 160 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
cut -c1-200 gpurun_out/r2_bench_n1.json gpurun_out/r2_bench_ref.json
