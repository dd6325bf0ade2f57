mkdir -p gpurun_out
timeout -s KILL 300 python bench.py --workload config2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02o_cfg2.json 2> gpurun_out/r02o_cfg2.err
echo "cfg2 rc=$?" > gpurun_out/r02o_env.log
timeout -s KILL 400 python bench.py --workload config4 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02o_cfg4.json 2> gpurun_out/r02o_cfg4.err
echo "cfg4 rc=$?" >> gpurun_out/r02o_env.log
timeout -s KILL 200 python -m pytest tests/test_gpu_path.py -m gpu -x -q -k "mixed_operators or cholesky or lu_fallback or wrappers" > gpurun_out/r02o_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02o_env.log
