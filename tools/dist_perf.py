"""[torchrun] tools/dist_perf.py N_TETRADS [steps random_mode kind laps lap_gap qfac scan_mode] : step time of the (sharded) engine.
The timed steps run as the engine runs them in production (one CUDA graph per step, in-engine barriers);
a second short pass with the per-phase timers on gives the breakdown."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
pkg = importlib.import_module("msc-project_b200"); dmod = importlib.import_module("msc-project_b200.dist")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
a = sys.argv[1:]
n = int(a[0]); steps = int(a[1]) if len(a) > 1 else 20; random_mode = int(a[2]) if len(a) > 2 else 1
kind = a[3] if len(a) > 3 else "ring"; laps = int(a[4]) if len(a) > 4 else 2; gap = float(a[5]) if len(a) > 5 else 40.0
qfac = float(a[6]) if len(a) > 6 else 0.0
scan_mode = int(a[7]) if len(a) > 7 else 4
s = pkg.make_ring(n, kind=kind, laps=laps, lap_gap=gap)
g = pkg.Engine(s, device=local)
g.set_nb_consts(100.0, qfac)
g.set_nb_scan_mode(scan_mode)
sh = dmod.ShardedEngine(dmod.EngineBackend(g, torch.device("cuda", local)), dist, s.n, rank, world)
sh.set_rng(1000000000, 1, 0); g.set_rng(1000000000, 1, 0)
if world == 1:
    sh.step = lambda k, mode: g.step(k, mode)      # one GPU: edmd_step (graph replay), the production entry point
t0 = time.time(); npairs = sh.build_pairlist(); tb = time.time() - t0
sh.step(3, random_mode); g.sync()
if dist: dist.barrier()
g.mark(0); t0 = time.perf_counter(); sh.step(steps, random_mode); t_host = time.perf_counter() - t0; g.mark(1)
ms = g.mark_elapsed(0, 1) / steps
t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
if dist: dist.all_reduce(t, op=dist.ReduceOp.MAX)
st = g.nb_stats()
g.timing(True); g.timing_reset()
sh.step(5, random_mode); g.sync()
ph = {nm: round(g.timing_get(i)[0] / 5, 4) for nm, i in (("ed", 0), ("rng", 3), ("scan", 1), ("finish", 2))}
g.timing(False)
t0 = time.time(); sh.merge(); npairs2 = sh.build_pairlist(); g.sync(); t_rebuild = time.time() - t0
if rank == 0:
    print(json.dumps(dict(n=n, world=world, kind=kind, laps=laps, gap=gap, qfac=qfac, random=random_mode, scan_mode=scan_mode, pairs=npairs,
                          first_build_s=round(tb, 3), merge_rebuild_ms=round(t_rebuild * 1e3, 2), ms_per_step=round(float(t), 4),
                          host_enqueue_ms=round(t_host / steps * 1e3, 4), phases_rank0=ph, **st)), flush=True)
sh.close()
if dist:
    dist.barrier(); dist.destroy_process_group()
