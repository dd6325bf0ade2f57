mkdir -p gpurun_out
timeout -s KILL 60 python tools/lu_bench.py 14336 > gpurun_out/r02t_lu.jsonl 2> gpurun_out/r02t_lu.err
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02t_lu_launches.csv python tools/lu_bench.py 14336 > gpurun_out/r02t_ncu.log 2>&1
echo "ncu rc=$?" > gpurun_out/r02t_env.log
