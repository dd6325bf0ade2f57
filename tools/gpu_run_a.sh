mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r02a_env.log 2>&1
nproc >> gpurun_out/r02a_env.log; free -g >> gpurun_out/r02a_env.log
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "dgemm" > gpurun_out/r02a_gemm.log 2>&1
echo "gemm rc=$?" >> gpurun_out/r02a_env.log
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02a_env.log
timeout -s KILL 200 python tools/syrk_bench.py 6656 3584 1024 > gpurun_out/r02a_syrk_tma.log 2>&1
PMR_GEMM_NO_TMA=1 timeout -s KILL 200 python tools/syrk_bench.py 6656 3584 1024 > gpurun_out/r02a_syrk_ring.log 2>&1
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?" >> gpurun_out/r02a_env.log
PMR_GEMM_NO_TMA=1 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-check > gpurun_out/r02a_bench_ring.json 2> gpurun_out/r02a_bench_ring.err
echo "bench ring rc=$?" >> gpurun_out/r02a_env.log
