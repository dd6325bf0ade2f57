mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -s KILL 200 $TR --master-port 29551 tools/allgather_probe.py > gpurun_out/r02m_probe.log 2>&1
timeout -s KILL 400 $TR --master-port 29552 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu --no-check > gpurun_out/r02m_n2.json 2> gpurun_out/r02m_n2.err
echo "n2 rc=$?" > gpurun_out/r02m_env.log
