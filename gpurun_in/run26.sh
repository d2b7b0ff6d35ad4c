timeout 85 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t26.log 2>&1; tail -4 gpurun_out/r2_t26.log
timeout 70 python bench.py --steps 50 --no-cpu-baseline > gpurun_out/bench_r02_mma_n1.json 2> gpurun_out/bench_r02_mma_n1.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r02_mma_n1.json').read().splitlines() if l.startswith('{')][-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['kernel_ms'], d['roofline']['kernel'], d['roofline']['frac'], d['extras'].get('c2'))" || tail -5 gpurun_out/bench_r02_mma_n1.err
