mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_path.py -m gpu -q -x -k "getrf or lu or pole or fallback or singular" > gpurun_out/r02s_tests.log 2>&1
echo "tests rc=$?" > gpurun_out/r02s_env.log
timeout -s KILL 150 python tools/lu_bench.py > gpurun_out/r02s_lu.jsonl 2> gpurun_out/r02s_lu.err
echo "lu rc=$?" >> gpurun_out/r02s_env.log
timeout -s KILL 60 python tools/lu_panel_phases.py 14336 > gpurun_out/r02u_phases_after.log 2>&1
