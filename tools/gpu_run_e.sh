mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/r02e_tests.log 2>&1
echo "tests rc=$?" > gpurun_out/r02e_env.log
timeout -s KILL 200 python tools/syrk_bench.py 6656 1024 > gpurun_out/r02e_syrk.log 2>&1
PMR_GEMM_TMA_REGS=112 timeout -s KILL 200 python tools/syrk_bench.py 6656 1024 > gpurun_out/r02e_syrk_r112.log 2>&1
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
echo "bench rc=$?" >> gpurun_out/r02e_env.log
PMR_GEMM_TMA_REGS=112 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-check > gpurun_out/r02e_bench_r112.json 2> gpurun_out/r02e_bench_r112.err
echo "bench r112 rc=$?" >> gpurun_out/r02e_env.log
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:potf2_trtri_kernel -s 3 -c 1 -f -o gpurun_out/r02e_potf2 python tools/potrf_only.py 2048 2 > gpurun_out/r02e_ncu_potf2.log 2>&1
echo "ncu potf2 rc=$?" >> gpurun_out/r02e_env.log
