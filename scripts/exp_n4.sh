mkdir -p gpurun_out
for S in 20 100; do
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 4 --steps $S > gpurun_out/exp_n4_steps$S.json 2> gpurun_out/exp_n4_steps$S.err
echo "steps $S rc=$?"; grep "bench +" gpurun_out/exp_n4_steps$S.err | tr '[' '\n' | grep "rank 0" | tail -5; cut -c1-120 gpurun_out/exp_n4_steps$S.json
done
