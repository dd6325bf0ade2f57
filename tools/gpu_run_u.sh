mkdir -p gpurun_out
timeout -s KILL 100 python tools/lu_panel_phases.py > gpurun_out/r02u_phases.log 2>&1
echo "rc=$?" > gpurun_out/r02u_env.log
