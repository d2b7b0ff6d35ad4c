timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t5.log 2>&1; tail -4 gpurun_out/r2_t5.log
P="python tools/dist_perf.py 100000 50 1"
echo default; timeout 200 $P 2>&1 | tail -1
echo sup1; EDMD_SUPERPOSE=1 timeout 200 $P 2>&1 | tail -1
echo nobranch; EDMD_RNG_BRANCH=0 timeout 200 $P 2>&1 | tail -1
for k in 2 3 6 8 16; do echo rngctas $k; EDMD_RNG_CTAS=$k timeout 200 $P 2>&1 | tail -1; done
echo n12500; timeout 100 python tools/dist_perf.py 12500 100 1 2>&1 | tail -1
echo n12500sup1; EDMD_SUPERPOSE=1 timeout 100 python tools/dist_perf.py 12500 100 1 2>&1 | tail -1
