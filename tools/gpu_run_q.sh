mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests/test_gpu_multi.py tests/test_gpu_td.py -m gpu -x -q > gpurun_out/r02q_tests.log 2>&1
echo "tests rc=$?" > gpurun_out/r02q_env.log
