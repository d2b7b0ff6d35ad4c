export EDMD_PEER_TIMEOUT_S=30
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_t4.log 2>&1; tail -5 gpurun_out/r2_t4.log
for N in 8 4 2; do
  timeout 600 $TR --nproc-per-node $N --master-port 2961$N bench.py --gpus $N > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err
  python -c "import json;d=json.load(open('gpurun_out/bench_r02_n$N.json'));print($N, d['ms_per_step'], d['e2e']['ms_per_step'], d.get('multi_gpu_parity'), d['kernel_ms'], d['ntsync_block'])" || tail -5 gpurun_out/bench_r02_n$N.err
done
timeout 600 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench_r02_n1b.json 2> gpurun_out/bench_r02_n1b.err
python -c "import json;d=json.load(open('gpurun_out/bench_r02_n1b.json'));print(1, d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'], d['ntsync_block'])" || tail -5 gpurun_out/bench_r02_n1b.err
timeout 300 $TR --nproc-per-node 8 --master-port 29631 tools/dist_perf.py 1000000 10 1 > gpurun_out/r2_1m_n8.log 2>&1; tail -1 gpurun_out/r2_1m_n8.log
