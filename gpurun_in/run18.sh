set -x
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/bench_r02_final_n1.err; tail -c 600 gpurun_out/bench_r02_final_n1.json; tail -3 gpurun_out/bench_r02_final_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_final_ref.json 2> gpurun_out/bench_r02_final_ref.err; tail -c 400 gpurun_out/bench_r02_final_ref.json
C3="python tools/prof_step.py 100000 ring 2 40 0 1 3"
COILQ="python tools/prof_step.py 10000 coil 2 40 83.016 0 3"
timeout 300 $C3 > gpurun_out/plain_c3.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r02_final_c3.csv $C3 > gpurun_out/ncu_l2.log 2>&1
timeout 300 $C3 > gpurun_out/plain_c3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"superpose2|ed_project|random_terms|nb_scan_blocks|finish_unmarked|group_bounds|pair_select" -s 9 -c 8 -o gpurun_out/prof_r02_final_c3 -f $C3 > gpurun_out/ncu_f4.log 2>&1
timeout 300 $COILQ > gpurun_out/plain_coil.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:"nb_force_kernel|nb_scan_cull|nb_scan_blocks" -s 3 -c 3 -o gpurun_out/prof_r02_final_force -f $COILQ > gpurun_out/ncu_f5.log 2>&1
tail -n 2 gpurun_out/ncu_f4.log; tail -n 2 gpurun_out/ncu_f5.log
