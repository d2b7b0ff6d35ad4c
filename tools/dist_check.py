"""torchrun --nproc-per-node N tools/dist_check.py : the sharded engine (peer memory over NVLink, in-engine
barriers, one CUDA graph per step) vs the single-GPU engine, bit for bit — coordinates, velocities, energies —
over steps, a merge + pair-list rebuild, and more steps, with Langevin random forces on."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
pkg = importlib.import_module("msc-project_b200")
dmod = importlib.import_module("msc-project_b200.dist")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ok = True
cases = (("coil60-replicated", pkg.make_ring(60, kind="coil", laps=2, lap_gap=9.0), True, 0.0),
         ("coil200-uneven", pkg.make_ring(200, kind="coil", laps=4, lap_gap=9.0), False, 0.0),
         ("coil200-qfac", pkg.make_ring(200, kind="coil", laps=4, lap_gap=9.0), False, 332.064 / 4.0),
         ("ring2000", pkg.make_ring(2000, kind="ring"), False, 0.0))
for name, s, replicate, qfac in cases:
    g = pkg.Engine(s, device=local)
    g.set_nb_consts(100.0, qfac)
    sh = dmod.ShardedEngine(dmod.EngineBackend(g, torch.device("cuda", local)), dist, s.n, rank, world, replicate=replicate)
    sh.set_rng(1000000000, 1, 0)
    npairs = sh.build_pairlist()
    sh.step(5, 1)
    sh.merge()
    npairs2 = sh.build_pairlist()
    sh.step(3, 1)
    sh.step(1, 1)
    sh.finish()
    c, v = g.get_state()
    en = g.energies()
    if rank == 0:
        g1 = pkg.Engine(s, device=local)
        g1.set_nb_consts(100.0, qfac)
        assert g1.build_pairlist() == npairs
        g1.set_rng(1000000000, 1, 0)
        g1.step(5, 1); g1.merge()
        assert g1.build_pairlist() == npairs2
        g1.step(4, 1)
        c1, v1 = g1.get_state()
        e1 = g1.energies()
        same = np.array_equal(c, c1) and np.array_equal(v, v1) and np.array_equal(en, e1)
        print(name, "world", world, "pairs", npairs, npairs2, "bitwise crd", np.array_equal(c, c1), "vel", np.array_equal(v, v1),
              "energies", np.array_equal(en, e1), "maxdiff", np.abs(c - c1).max(), flush=True)
        ok = ok and same
        g1.close()
    sh.close()
    g.close()
    dist.barrier()
flag = torch.tensor([1 if ok else 0], device=f"cuda:{local}")
dist.broadcast(flag, src=0)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
