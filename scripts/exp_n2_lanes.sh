mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -s -k "multi_gpu" 2>&1 | tail -8
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 20 > gpurun_out/exp_n2.json 2> gpurun_out/exp_n2.err
echo "rc=$?"; grep "bench +" gpurun_out/exp_n2.err | tr '[' '\n' | grep "rank 0" | tail -8; cut -c1-150 gpurun_out/exp_n2.json
