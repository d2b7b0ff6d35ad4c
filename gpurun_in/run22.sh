tools/micro/dmma_rate > gpurun_out/dmma_rate.log 2>&1; cat gpurun_out/dmma_rate.log
timeout 150 python tools/ed_mma_check.py 100000 > gpurun_out/ed_mma_check.log 2>&1; cut -c1-330 gpurun_out/ed_mma_check.log
