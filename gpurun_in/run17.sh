timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t17.log 2>&1; tail -3 gpurun_out/r2_t17.log
echo c3; timeout 200 python tools/dist_perf.py 100000 50 1 2>&1 | tail -1
echo n12500; timeout 100 python tools/dist_perf.py 12500 100 1 2>&1 | tail -1
