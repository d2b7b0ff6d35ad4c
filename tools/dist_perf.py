"""torchrun tools/dist_perf.py N_TETRADS [mode] : step time of the sharded engine (mode: halo|allgather|replicate)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
pkg = importlib.import_module("msc-project_b200"); dmod = importlib.import_module("msc-project_b200.dist")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]); mode = sys.argv[2] if len(sys.argv) > 2 else "halo"; steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
random_mode = int(sys.argv[4]) if len(sys.argv) > 4 else 0
scan_mode = int(sys.argv[5]) if len(sys.argv) > 5 else 1
s = pkg.make_ring(n, kind="ring")
g = pkg.Engine(s, device=local)
g.set_nb_scan_mode(scan_mode)
sh = dmod.ShardedEngine(dmod.EngineBackend(g, torch.device("cuda", local)), dist if world > 1 else None, s.n, rank, world,
                        replicate=(mode == "replicate"), halo=(mode == "halo"))
t0 = time.time(); npairs = sh.build_pairlist(); tb = time.time() - t0
sh.step(3, random_mode); g.sync()
if world > 1: dist.barrier()
g.timing(True); g.timing_reset()
g.mark(0); sh.step(steps, random_mode); g.mark(1)
ms = g.mark_elapsed(0, 1) / steps
t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
k = {nm: (lambda r: r[0] / r[1] if r[1] else None)(g.timing_get(i)) for nm, i in (("ed", 0), ("scan", 1), ("finish", 2), ("rng", 3))}
if rank == 0:
    halo = sh.halo_plan.recv_idx.size if sh.halo_plan is not None else None
    print(f"n={n} world={world} mode={mode} scan_mode={scan_mode} random={random_mode} pairs={npairs} build={tb:.2f}s step={float(t):.3f} ms  kernels(ms/launch)={k} halo_rows={halo}", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
