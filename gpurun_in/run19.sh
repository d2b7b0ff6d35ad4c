export EDMD_PEER_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_t19.log 2>&1; tail -3 gpurun_out/r2_t19.log
for N in 8 4 2; do
  timeout 600 $TR --nproc-per-node $N --master-port 2991$N bench.py --gpus $N > gpurun_out/bench_r02_final_n$N.json 2> gpurun_out/bench_r02_final_n$N.err
  python -c "
import json
txt=open('gpurun_out/bench_r02_final_n$N.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print($N, d['ms_per_step'], d['e2e']['ms_per_step'], d.get('multi_gpu_parity'), d['kernel_ms'], d['ntsync_block'])" || tail -5 gpurun_out/bench_r02_final_n$N.err
done
timeout 300 $TR --nproc-per-node 8 --master-port 29931 tools/dist_perf.py 1000000 10 1 > gpurun_out/r2_1m_n8_final.log 2>&1; tail -1 gpurun_out/r2_1m_n8_final.log
