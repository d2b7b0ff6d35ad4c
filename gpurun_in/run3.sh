set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t3.log 2>&1; tail -5 gpurun_out/r2_t3.log
timeout 900 python bench.py > gpurun_out/bench_r02_n1.json 2> gpurun_out/bench_r02_n1.err; tail -c 3000 gpurun_out/bench_r02_n1.json; tail -5 gpurun_out/bench_r02_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err; cat gpurun_out/bench_r02_ref.json
C3="python tools/prof_step.py 100000 ring 2 40 0 1 3"
COIL="python tools/prof_step.py 10000 coil 2 40 83.016 0 3"
timeout 300 $C3 > gpurun_out/plain_c3.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02_c3.csv $C3 > gpurun_out/ncu_l1.log 2>&1
timeout 300 $C3 > gpurun_out/plain_c3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"superpose|ed_project|random_terms|nb_scan_cull|finish_kernel|group_bounds|pair_select" -s 9 -c 8 -o gpurun_out/prof_r02_c3 -f $C3 > gpurun_out/ncu_f1.log 2>&1
timeout 300 $COIL > gpurun_out/plain_coil.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nb_force_kernel|nb_scan_cull" -s 2 -c 4 -o gpurun_out/prof_r02_force -f $COIL > gpurun_out/ncu_f2.log 2>&1
tail -3 gpurun_out/ncu_f1.log gpurun_out/ncu_f2.log gpurun_out/ncu_l1.log
