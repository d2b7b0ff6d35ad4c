export EDMD_PEER_TIMEOUT_S=60
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29941 tools/dist_perf.py 1000000 10 1 ring 2 40 0 2 > gpurun_out/c4_fp32_n8.log 2>&1; tail -1 gpurun_out/c4_fp32_n8.log
