timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t14.log 2>&1; tail -4 gpurun_out/r2_t14.log
for d in 9 1 65 4; do echo dense $d; EDMD_CULL_DENSE=$d timeout 200 python tools/dist_perf.py 100000 30 1 2>&1 | tail -1; EDMD_CULL_DENSE=$d timeout 100 python tools/dist_perf.py 10000 50 0 coil 2 40 2>&1 | tail -1;  EDMD_CULL_DENSE=$d timeout 100 python tools/dist_perf.py 10000 30 0 coil 2 40 83.016 2>&1 | tail -1; done
EDMD_CULL_DENSE=65 timeout 900 python -m pytest tests/test_gpu_edge.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
echo n12500; timeout 100 python tools/dist_perf.py 12500 100 1 2>&1 | tail -1
