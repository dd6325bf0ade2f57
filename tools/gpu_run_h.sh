mkdir -p gpurun_out
timeout -s KILL 120 python tools/potf2_phases.py > gpurun_out/r02h_potf2_phases.log 2>&1
echo "phases rc=$?" > gpurun_out/r02h_env.log
timeout -s KILL 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "potrf or potri or potrs" > gpurun_out/r02h_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r02h_env.log
timeout -s KILL 120 python tools/stage_times.py > gpurun_out/r02h_stage.log 2>&1
