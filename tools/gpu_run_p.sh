mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py tests/test_gpu_fullsize.py -m gpu -x -q > gpurun_out/r02p_tests.log 2>&1
echo "tests rc=$?" > gpurun_out/r02p_env.log
timeout -s KILL 400 python bench.py --workload config4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/r02p_cfg4.json 2> gpurun_out/r02p_cfg4.err
echo "cfg4 rc=$?" >> gpurun_out/r02p_env.log
