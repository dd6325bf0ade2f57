mkdir -p gpurun_out
(nproc; free -g | head -2; cat /sys/fs/cgroup/memory.max 2>/dev/null) > gpurun_out/r02j_env.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -s KILL 700 $TR --master-port 29531 bench.py --gpus 2 --workload config4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02j_n2_cfg4.json 2> gpurun_out/r02j_n2_cfg4.err
echo "n2 config4 rc=$?" >> gpurun_out/r02j_env.log
timeout -s KILL 600 $TR --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02j_n2.json 2> gpurun_out/r02j_n2.err
echo "n2 rc=$?" >> gpurun_out/r02j_env.log
grep "^\[bench" gpurun_out/r02j_n2_cfg4.err | tail -12
