timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t8.log 2>&1; tail -4 gpurun_out/r2_t8.log
timeout 200 python tools/dist_perf.py 100000 50 1 2>&1 | tail -1
timeout 100 python tools/dist_perf.py 12500 100 1 2>&1 | tail -1
timeout 100 python tools/dist_perf.py 1000 200 0 2>&1 | tail -1
C3="python tools/prof_step.py 100000 ring 2 40 0 1 3"
timeout 300 $C3 > gpurun_out/plain_c3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"superpose2|nb_scan_cull|group_bounds" -s 3 -c 3 -o gpurun_out/prof_r02_sup2 -f $C3 > gpurun_out/ncu_f3.log 2>&1
tail -n 3 gpurun_out/ncu_f3.log
