timeout 150 python tools/ed_mma_check.py 100000 > gpurun_out/ed_mma_check_v2.log 2>&1; cut -c1-330 gpurun_out/ed_mma_check_v2.log | tail -12
export EDMD_ED_MMA=1
timeout 100 ncu --set full --clock-control none --import-source on -k regex:ed_project_mma -c 1 -o gpurun_out/prof_r02_ed_mma -f python tools/prof_step.py 100000 ring 2 40 0 1 1 > gpurun_out/ncu_mma.log 2>&1; tail -3 gpurun_out/ncu_mma.log
