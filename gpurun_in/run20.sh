timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t20.log 2>&1; tail -3 gpurun_out/r2_t20.log
timeout 300 python bench.py --steps 50 --no-extras --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_quick.json').read().splitlines() if l.startswith('{')][-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['ntsync_block'], d['roofline']['kernel'], d['roofline']['frac'])" || tail -5 gpurun_out/bench_quick.err
timeout 100 python tools/dist_perf.py 10000 30 0 coil 2 40 83.016 2>&1 | tail -1
