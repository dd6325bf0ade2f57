mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_path.py -m gpu -x -q > gpurun_out/r02n_tests.log 2>&1
echo "tests rc=$?" > gpurun_out/r02n_env.log
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-e2e > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
echo "bench rc=$?" >> gpurun_out/r02n_env.log
