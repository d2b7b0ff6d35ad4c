timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t16.log 2>&1; tail -4 gpurun_out/r2_t16.log
echo c3; timeout 200 python tools/dist_perf.py 100000 50 1 2>&1 | tail -1
echo n12500; timeout 100 python tools/dist_perf.py 12500 100 1 2>&1 | tail -1
echo n25000; timeout 100 python tools/dist_perf.py 25000 100 1 2>&1 | tail -1
echo coil; timeout 100 python tools/dist_perf.py 10000 50 0 coil 2 40 2>&1 | tail -1
echo c1; timeout 100 python tools/dist_perf.py 1000 200 0 2>&1 | tail -1
