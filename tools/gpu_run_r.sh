mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -s KILL 300 $TR --master-port 29561 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/r02r_n2.json 2> gpurun_out/r02r_n2.err
echo "n2 rc=$?" > gpurun_out/r02r_env.log
