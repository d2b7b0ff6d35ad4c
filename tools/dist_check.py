"""torchrun --nproc-per-node N tools/dist_check.py : sharded engine (NCCL) vs single-GPU engine, bitwise."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
pkg = importlib.import_module("msc-project_b200")
dmod = importlib.import_module("msc-project_b200.dist")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
for name, s in (("coil64", pkg.make_ring(64 * 10 // 10 if False else 60, kind="coil", laps=2, lap_gap=9.0)),
                ("coil200-uneven", pkg.make_ring(200 if world != 3 else 200, kind="coil", laps=4, lap_gap=9.0))):
    g = pkg.Engine(s, device=local)
    sh = dmod.ShardedEngine(dmod.EngineBackend(g, torch.device("cuda", local)), dist, s.n, rank, world,
                            replicate=(name == "coil64"), halo=True)
    npairs = sh.build_pairlist()
    sh.set_rng(1000000000, 1, 0)
    sh.step(5, 1)
    sh.merge()
    sh.step(2, 1)
    sh.finish()
    c, v = g.get_state()
    if rank == 0:
        g1 = pkg.Engine(s, device=local)
        assert g1.build_pairlist() == npairs
        g1.set_rng(1000000000, 1, 0)
        g1.step(5, 1); g1.merge(); g1.step(2, 1)
        c1, v1 = g1.get_state()
        print(name, "world", world, "pairs", npairs, "bitwise crd", np.array_equal(c, c1), "vel", np.array_equal(v, v1),
              "maxdiff", np.abs(c - c1).max(), flush=True)
        assert np.array_equal(c, c1) and np.array_equal(v, v1)
    dist.barrier()
dist.destroy_process_group()
