timeout 150 python tools/ed_mma_check.py 100000 > gpurun_out/ed_mma_check_v3.log 2>&1; cut -c1-330 gpurun_out/ed_mma_check_v3.log | grep -v '"kernel": 2, "n"' | tail -14
