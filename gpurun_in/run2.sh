export EDMD_PEER_TIMEOUT_S=20
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t2.log 2>&1; tail -15 gpurun_out/r2_t2.log
timeout 300 python tools/dist_perf.py 100000 20 1 > gpurun_out/r2_c3_n1.log 2>&1; tail -2 gpurun_out/r2_c3_n1.log
timeout 300 $TR --master-port 29611 tools/dist_perf.py 100000 20 1 > gpurun_out/r2_c3_n2.log 2>&1; tail -2 gpurun_out/r2_c3_n2.log
timeout 200 python tools/dist_perf.py 10000 20 0 coil 2 40 > gpurun_out/r2_coil10k.log 2>&1; tail -1 gpurun_out/r2_coil10k.log
timeout 200 python tools/dist_perf.py 10000 20 0 ring > gpurun_out/r2_ring10k.log 2>&1; tail -1 gpurun_out/r2_ring10k.log
timeout 200 python tools/dist_perf.py 10000 20 0 coil 2 40 83.016 > gpurun_out/r2_coil10k_q.log 2>&1; tail -1 gpurun_out/r2_coil10k_q.log
timeout 200 python tools/dist_perf.py 10000 20 0 ring 2 40 83.016 > gpurun_out/r2_ring10k_q.log 2>&1; tail -1 gpurun_out/r2_ring10k_q.log
timeout 300 python tools/dist_perf.py 100000 20 1 coil 2 40 > gpurun_out/r2_coil100k.log 2>&1; tail -1 gpurun_out/r2_coil100k.log
timeout 200 $TR --master-port 29612 tools/dist_perf.py 1000 50 0 > gpurun_out/r2_1k_n2.log 2>&1; tail -1 gpurun_out/r2_1k_n2.log
timeout 200 python tools/dist_perf.py 1000 50 0 > gpurun_out/r2_1k_n1.log 2>&1; tail -1 gpurun_out/r2_1k_n1.log
