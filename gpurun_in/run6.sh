P="python tools/dist_perf.py 100000 50 1"
echo default; timeout 200 $P 2>&1 | tail -1
echo nobranch; EDMD_RNG_BRANCH=0 timeout 200 $P 2>&1 | tail -1
for k in 1 2 3 6 8 12 16; do echo rngctas $k; EDMD_RNG_CTAS=$k timeout 200 $P 2>&1 | tail -1; done
echo n12500; timeout 100 python tools/dist_perf.py 12500 100 1 2>&1 | tail -1
echo n12500nobranch; EDMD_RNG_BRANCH=0 timeout 100 python tools/dist_perf.py 12500 100 1 2>&1 | tail -1
