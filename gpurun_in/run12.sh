for pad in 0 9000 18000 27000 36000; do echo pad $pad; EDMD_SUP2_PAD=$pad timeout 200 python tools/dist_perf.py 100000 30 1 2>&1 | tail -1; done
