mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02c_tests.log 2>&1
echo "tests rc=$?" > gpurun_out/r02c_env.log
timeout -s KILL 200 python tools/syrk_bench.py 6656 3584 1024 > gpurun_out/r02c_syrk.log 2>&1
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
echo "bench rc=$?" >> gpurun_out/r02c_env.log
PMR_GEMM_TMA_INPLACE=0 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-check > gpurun_out/r02c_bench_noinplace.json 2> gpurun_out/r02c_bench_noinplace.err
echo "bench noinplace rc=$?" >> gpurun_out/r02c_env.log
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:dgemm_tma_kernel -s 1 -c 1 -f -o gpurun_out/r02c_trailing python tools/profile_kernels.py trailing > gpurun_out/r02c_ncu_trailing.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/r02c_env.log
timeout -s KILL 300 python tools/lu_bench.py > gpurun_out/r02c_lu.log 2>&1
echo "lu rc=$?" >> gpurun_out/r02c_env.log
