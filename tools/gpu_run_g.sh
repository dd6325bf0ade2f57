mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/r02g_tests.log 2>&1
echo "tests rc=$?" > gpurun_out/r02g_env.log
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 --no-e2e > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err
echo "bench rc=$?" >> gpurun_out/r02g_env.log
timeout -s KILL 120 python tools/potrf_only.py 7168 4 > gpurun_out/r02g_potrf.log 2>&1
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:potf2 -c 60 --csv --log-file gpurun_out/r02g_potf2_times.csv python tools/potrf_only.py 7168 4 > /dev/null 2>&1
echo "ncu rc=$?" >> gpurun_out/r02g_env.log
