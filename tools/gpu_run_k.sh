mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q --durations=3 > gpurun_out/r02k_tests.log 2>&1
echo "tests rc=$?" > gpurun_out/r02k_env.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/r02k_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02k_env.log
timeout -s KILL 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
echo "bench rc=$?" >> gpurun_out/r02k_env.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02k_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-check > gpurun_out/r02k_ncu_bench.log 2>&1
echo "ncu launches rc=$?" >> gpurun_out/r02k_env.log
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:dgemm_tma_kernel -s 1 -c 1 -f -o gpurun_out/r02k_trailing python tools/profile_kernels.py trailing > gpurun_out/r02k_ncu_trailing.log 2>&1
echo "ncu full rc=$?" >> gpurun_out/r02k_env.log
