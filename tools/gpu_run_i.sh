mkdir -p gpurun_out
for v in 4 3; do PMR_GEMM_TMA_STAGES=$v timeout -s KILL 150 python tools/syrk_bench.py 6656 1024 > gpurun_out/r02i_syrk_st$v.log 2>&1; done
PMR_GEMM_TILE=128 timeout -s KILL 150 python tools/syrk_bench.py 6656 1024 > gpurun_out/r02i_syrk_tile128.log 2>&1
timeout -s KILL 600 python bench.py --steps 5 --warmup 3 --no-e2e > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err
echo "bench rc=$?" > gpurun_out/r02i_env.log
PMR_GEMM_TMA_STAGES=3 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-check > gpurun_out/r02i_bench_st3.json 2> gpurun_out/r02i_bench_st3.err
PMR_GEMM_TILE=128 timeout -s KILL 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu --no-check > gpurun_out/r02i_bench_tile128.json 2> gpurun_out/r02i_bench_tile128.err
