mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02b_tests.log 2>&1
echo "tests rc=$?" > gpurun_out/r02b_env.log
PMR_GEMM_C_LATE=1 timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "dgemm or potrf" > gpurun_out/r02b_tests_clate.log 2>&1
echo "tests clate rc=$?" >> gpurun_out/r02b_env.log
timeout -s KILL 200 python tools/syrk_bench.py 6656 3584 1024 > gpurun_out/r02b_syrk.log 2>&1
PMR_GEMM_C_LATE=1 timeout -s KILL 200 python tools/syrk_bench.py 6656 3584 1024 > gpurun_out/r02b_syrk_clate.log 2>&1
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
echo "bench rc=$?" >> gpurun_out/r02b_env.log
PMR_GEMM_C_LATE=1 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-check > gpurun_out/r02b_bench_clate.json 2> gpurun_out/r02b_bench_clate.err
echo "bench clate rc=$?" >> gpurun_out/r02b_env.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-check > gpurun_out/r02b_ncu_bench.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/r02b_env.log
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:dgemm_tma_kernel -s 1 -c 1 -f -o gpurun_out/r02b_trailing python tools/profile_kernels.py trailing > gpurun_out/r02b_ncu_trailing.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/r02b_env.log
