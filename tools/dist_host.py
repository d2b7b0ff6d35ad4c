"""torchrun tools/dist_host.py N_TETRADS : host enqueue time vs device time per sharded step (halo mode, random on)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
pkg = importlib.import_module("msc-project_b200"); dmod = importlib.import_module("msc-project_b200.dist")
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]); steps = 40
s = pkg.make_ring(n, kind="ring")
g = pkg.Engine(s, device=local)
sh = dmod.ShardedEngine(dmod.EngineBackend(g, torch.device("cuda", local)), dist if world > 1 else None, s.n, rank, world, halo=True)
npairs = sh.build_pairlist(); sh.set_rng(1000000000, 1, 0)
sh.step(5, 1); g.sync()
if world > 1: dist.barrier()
torch.cuda.synchronize()
g.mark(0); t0 = time.perf_counter(); sh.step(steps, 1); t_host = time.perf_counter() - t0; g.mark(1)
ms = g.mark_elapsed(0, 1) / steps
print(f"rank {rank}/{world}: n={n} pairs={npairs} device {ms:.3f} ms/step, host enqueue {t_host / steps * 1e3:.3f} ms/step", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
