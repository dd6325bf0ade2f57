mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout -s KILL 600 $TR --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02d_n2.json 2> gpurun_out/r02d_n2.err
echo "n2 rows rc=$?" > gpurun_out/r02d_env.log
timeout -s KILL 400 $TR --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --hessian replicated --no-e2e --no-cpu --no-check > gpurun_out/r02d_n2_repl.json 2> gpurun_out/r02d_n2_repl.err
echo "n2 replicated rc=$?" >> gpurun_out/r02d_env.log
timeout -s KILL 600 $TR --master-port 29513 bench.py --gpus 2 --workload config4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02d_n2_cfg4.json 2> gpurun_out/r02d_n2_cfg4.err
echo "n2 config4 rc=$?" >> gpurun_out/r02d_env.log
timeout -s KILL 300 $TR --master-port 29514 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r02d_n2_ref.json 2> gpurun_out/r02d_n2_ref.err
echo "n2 reference rc=$?" >> gpurun_out/r02d_env.log
tail -c 600 gpurun_out/r02d_n2.err
