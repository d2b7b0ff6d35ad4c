export CHECK_PREFETCH_MODES=1,0,2
timeout 150 python tools/ed_mma_check.py 100000 > gpurun_out/ed_mma_check_v4.log 2>&1; cut -c1-330 gpurun_out/ed_mma_check_v4.log | grep -v '"vs_oracle"' | tail -12
