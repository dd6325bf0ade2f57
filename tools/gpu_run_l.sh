mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout -s KILL 700 $TR --master-port 29541 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02l_n4.json 2> gpurun_out/r02l_n4.err
echo "n4 rc=$?" > gpurun_out/r02l_env.log
grep "^\[bench" gpurun_out/r02l_n4.err | tail -8
