mkdir -p gpurun_out
(nproc; free -g | head -2; nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8; cat /sys/fs/cgroup/memory.max 2>/dev/null) > gpurun_out/r02f_env.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout -s KILL 840 $TR --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02f_n8.json 2> gpurun_out/r02f_n8.err
echo "n8 rc=$?" >> gpurun_out/r02f_env.log
grep "^\[bench" gpurun_out/r02f_n8.err | tail -30
