#!/usr/bin/env python
"""bench.py — EDMD hot-path throughput on B200 (contract: see the task statement).

Workload (BASELINE.json configs[3], "C3" — the configuration the metric quotes at 1/2/4/8 B200): synthetic
100,000-tetrad replicated GC DNA ring, cutoff pair list (mole_Cutoff 30 A, cell-list builder), Langevin
random forces ON with the reference's rand() stream.  A "step" = one pass of the hot path:
ED forces -> random terms -> NB tetrad-pair forces (+clip) -> velocity + coordinate update.
Metric: tetrad-pair interactions per second (steps/s alongside).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

N > 1: launched by torchrun, one rank per GPU; the SAME workload is sharded (strong scaling): every rank owns a
contiguous block of tetrads and a slice of the pair list; the per-step exchanges (halo pull, mask broadcast,
barriers) are kernels inside the engine's CUDA graph over NVLink peer memory (msc-project_b200/csrc/peer.cu).

--impl reference: the reference's own CPU implementation of the path (oracle/_ref — the unmodified reference
sources compiled here — else the plain-C port) on all host threads, rank 0, on a bounded sample of C3.

Extras beside the headline (N = 1): C1 (configs[1]: 1,000 tetrads, all-pairs list, random off — round 1's
headline) with the brute-force FP64 distance kernel's roofline, and a contact-bearing coil (10,000 tetrads,
~15 % of them in steric contact) with and without electrostatics, where the exact NB force kernel is timed.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_C3 = 100000
ATOMS = 252
FLOP_PER_ATOM_PAIR = 9          # 3 SUB + 3 MUL + 2 ADD + 1 MAX per atom pair
METRIC = "tetrad_pair_interactions_per_sec"
UNIT = "tetrad-pairs/s"
WORKLOAD = ("C3: synthetic 100000-tetrad replicated GC DNA ring, cutoff pair list (mole_Cutoff 30 A, cell-list "
            "builder), Langevin random forces on (reference rand() stream)")
RNG_TIME0 = 1000000000
# real (de-duplicated) DRAM bytes per tetrad of the per-tetrad kernels, S = 256 atom slots, FP64:
#   superpose: read crd (3*256*8) + write 24 doubles;  project: read crd + 24 doubles, write crd + fed + e_ed;
#   random: write rnd;  finish: read vel, fed, rnd, crd, write vel, crd2 (parameter sets are L2-resident)
PLANE = 3 * 256 * 8
BYTES = {"superpose": PLANE + 192, "project": 3 * PLANE + 200, "random": PLANE, "finish": 6 * PLANE + 8}


def load_pkg():
    return importlib.import_module("msc-project_b200")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []
        self.t_begin = self.t_end = None

    def mark_begin(self):
        self.t_begin = time.time()

    def mark_end(self):
        self.t_end = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t_wait = time.time()
            while not self.lines and time.time() - t_wait < 3.0:   # nvidia-smi needs ~0.1-1 s for its first line
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        import datetime
        all_sm, all_reasons = [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                v_sm = float(f[1]); mx = float(f[2])
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
            except ValueError:
                continue
            mine = {nm for nm, v in zip(names, f[5:9]) if v.lower().startswith("active")}
            all_sm.append(v_sm); all_reasons |= mine
            # keep the samples taken inside the timed region (50 ms slack for the sampling period)
            if self.t_begin and self.t_end and not (self.t_begin - 0.05 <= ts <= self.t_end + 0.05):
                continue
            sm.append(v_sm); reasons |= mine
        window = "timed region"
        if len(sm) < 3:     # a timed region shorter than a few sampling periods: use every sample of the run
            sm, reasons, window = all_sm, all_reasons, "whole run (warm-up, timed region and variants, all under load)"
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


def bind_near_gpu(local):
    """Run this rank (and first-touch its pinned host buffers) on the CPUs of the NUMA node the GPU hangs off:
    host<->device copies that cross the socket interconnect run at a fraction of the PCIe rate.  Best effort."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id if hasattr(torch.cuda.get_device_properties(local), "pci_bus_id") else None
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(local), "pci_device_id", 0)
        if bus is None:
            return None
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/local_cpulist"
        txt = open(path).read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return txt
    except Exception:
        pass
    return None


def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def ring_pair_count(s):
    """Length of the reference's pair list (master.cpp:180-210) for a planar ring, without building it.  Same
    predicate in the same arithmetic (centroid summed sequentially in atom order, master.cpp:157-176; no FMA); only
    chain neighbours — small |i-j| or |i-j| close to n across the ring closure — can be inside mole_Cutoff."""
    import numpy as np
    n = s.n
    c = s.crd.reshape(n, ATOMS, 3)
    cen = np.cumsum(c, axis=1)[:, -1, :] / ATOMS              # cumsum adds strictly left to right
    cut2, least = float(s.params[5]) ** 2, float(s.params[7])
    seps = sorted(set(range(1, 64)) | set(range(max(1, n - 63), n)))
    total = 0
    for sep in seps:
        if not (sep > least and sep < n - least):
            continue
        d = cen[:n - sep] - cen[sep:]
        r = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        total += int((r < cut2).sum())
    return total


def config_dict(npairs):
    """The workload description both arms print (identical dicts: the driver compares them)."""
    return {"workload": WORKLOAD, "tetrads": N_C3, "tetrad_pairs": int(npairs), "atoms_per_tetrad": ATOMS,
            "atom_pairs_per_step": int(npairs) * ATOMS * ATOMS, "random_forces": "on",
            "inputs": "state planes 3.1 GB per GPU (> 126 MB L2): resident in HBM, nothing to flush between steps"}


def cpu_variants(pyoracle, o, full, npairs_full, m, np_sub, cores, kind, t_ed, t_rng, t_int):
    """BASELINE.md section 3's other two CPU figures, each on a bounded sample: the same build on ONE core, and the build
    at the optimisation level the reference ships with (-O0: its Makefile leaves CFLAGS empty) on all cores."""
    out = {}
    try:
        p1 = max(1, np_sub // cores)                       # one thread's share of the sample's pair list
        o.build_pairs()
        t_nb1 = o.time_nb(0, p1, 1)
        o.set_pairs([])
        step1 = (t_ed + t_rng + t_int) * (full.n / m) + t_nb1 * (npairs_full / p1)
        out["one_core"] = {"value": npairs_full / step1, "unit": UNIT, "cores": 1, "kind": kind, "ms_per_step": step1 * 1e3,
                           "sample": f"NB on the first {p1} tetrad pairs of the sample, 1 thread: {t_nb1:.3f} s; ED, random terms "
                                     f"and update as measured above, not divided"}
    except Exception as ex:                                # a variant must never cost the bench line
        out["one_core"] = {"error": repr(ex)}
    try:
        k0 = "reference_O0"
        if kind == "reference" and pyoracle.available(k0):
            m0 = max(200, m // 8)
            sub0 = full.subset(range(m0))
            o0 = pyoracle.Oracle(sub0, k0)
            o0.set_time(RNG_TIME0)
            np0 = o0.build_pairs()
            t_nb0 = o0.time_nb(0, np0, cores)
            o0.set_pairs([])
            t0 = time.perf_counter(); o0.ed_forces(); t_ed0 = time.perf_counter() - t0
            t0 = time.perf_counter(); o0.random_terms(); t_rng0 = time.perf_counter() - t0
            t0 = time.perf_counter(); o0.update_velocities(); o0.update_coordinates(); t_int0 = time.perf_counter() - t0
            o0.close()
            step0 = ((t_ed0 + t_rng0) / cores + t_int0) * (full.n / m0) + t_nb0 * (npairs_full / np0)
            out["shipped_O0"] = {"value": npairs_full / step0, "unit": UNIT, "cores": cores, "kind": "reference, g++ -O0 (the "
                                 "reference's Makefile sets no optimisation flag)", "ms_per_step": step0 * 1e3,
                                 "sample": f"first {m0} tetrads ({np0} pairs): NB {t_nb0:.3f} s on {cores} threads, ED {t_ed0:.3f} s, "
                                           f"random terms {t_rng0:.3f} s (1 thread, counted /{cores}), update {t_int0:.3f} s serial"}
        else:
            out["shipped_O0"] = {"unavailable": "oracle/_ref/libedmd_ref_O0.so is built only where /root/reference exists"}
    except Exception as ex:
        out["shipped_O0"] = {"error": repr(ex)}
    return out


def cpu_sample(full, npairs_full, m, kind_wanted=None, reps=2, cache=None, with_variants=False):
    """The reference CPU path on a bounded sample of C3: the first `m` tetrads of the ring (a chain segment with
    its own cutoff pair list), all host threads.  Per-tetrad and per-pair costs are uniform along the ring, so the
    full step is assembled as  t_tetrad_part * n/m + t_nb * pairs/pairs_sample.  The step is modelled the way the
    reference's MPI run with W = cores workers executes it: ED and random terms in parallel worker blocks
    (worker.cpp:155-162; measured on one thread and divided by W — libc rand() is per process there), NB in
    parallel pair blocks (worker.cpp:166-179, measured on W threads), velocity/coordinate update serial on the
    master (master.cpp:327-343)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import pyoracle
    kind = kind_wanted or ("reference" if pyoracle.available("reference") else "port")
    if not pyoracle.available(kind):
        pyoracle.build(kind)
    cores = host_threads()
    if cache is not None and "o" in cache:          # --impl reference: the sample system is built once, timed per step
        o, sub, crd0, vel0 = cache["o"], cache["sub"], cache["crd0"], cache["vel0"]
        o.set_state(crd0, vel0)
    else:
        sub = full.subset(range(m))
        o = pyoracle.Oracle(sub, kind)
        crd0, vel0 = sub.crd.copy(), sub.vel.copy()
        if cache is not None:
            cache.update(o=o, sub=sub, crd0=crd0, vel0=vel0)
    o.set_time(RNG_TIME0)
    np_sub = o.build_pairs()
    t_nb = min(o.time_nb(0, np_sub, cores) for _ in range(reps))
    o.set_pairs([])
    t0 = time.perf_counter(); o.ed_forces(); t_ed = time.perf_counter() - t0
    t0 = time.perf_counter(); o.random_terms(); t_rng = time.perf_counter() - t0
    t0 = time.perf_counter(); o.update_velocities(); o.update_coordinates(); t_int = time.perf_counter() - t0
    scale_t, scale_p = full.n / m, npairs_full / np_sub
    step_s = ((t_ed + t_rng) / cores + t_int) * scale_t + t_nb * scale_p
    variants = None
    if with_variants:
        variants = cpu_variants(pyoracle, o, full, npairs_full, m, np_sub, cores, kind, t_ed, t_rng, t_int)
    if cache is None:
        o.close()
    return {"value": npairs_full / step_s, "unit": UNIT, "cores": cores, "kind": kind, "variants": variants,
            "sample": f"first {m} of {full.n} tetrads ({np_sub} of {npairs_full} tetrad pairs): NB {t_nb:.3f} s on {cores} threads; "
                      f"ED {t_ed:.3f} s and random terms {t_rng:.3f} s measured on 1 thread and counted /{cores} (perfectly "
                      f"parallel worker blocks, one libc stream per MPI worker); velocity+coordinate update {t_int:.3f} s serial "
                      f"(master); scaled by tetrads and pairs",
            "steps_per_s": 1.0 / step_s, "ms_per_step": step_s * 1e3}


def run_reference(args, rank):
    if rank != 0:
        return
    pkg = load_pkg()
    s = pkg.make_ring(N_C3, kind="ring")
    npairs = ring_pair_count(s)
    m = args.cpu_sample
    vals, last, cache = [], None, {}
    for it in range(args.warmup + args.steps):
        r = cpu_sample(s, npairs, m, reps=1, cache=cache)
        if it >= args.warmup:
            vals.append(r); last = r
    v = statistics.mean(x["value"] for x in vals)
    sps = statistics.mean(x["steps_per_s"] for x in vals)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / sps, "steps_per_sec": sps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(npairs),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": last["cores"], "kind": last["kind"],
                             "sample": last["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def phase_times(g, step_fn, reps):
    """Per-phase device times (ms per step) from a short pass with the engine's CUDA-event timers on."""
    g.timing(True); g.timing_reset()
    step_fn(reps)
    g.sync()
    out = {}
    for nm, idx in (("ed", g.T_ED), ("superpose", g.T_SUP), ("random", g.T_RNG), ("nb_distance", g.T_SCAN), ("finish", g.T_ACC)):
        ms, _ = g.timing_get(idx)
        out[nm] = ms / reps
    out["project"] = out["ed"] - out["superpose"]
    g.timing(False)
    return out


def timed_steps(g, step_fn, k, slot):
    g.mark(slot); step_fn(k); g.mark(slot + 1)
    return g.mark_elapsed(slot, slot + 1) / k


def extra_c1(pkg, eng_mod, local, peak_tf):
    """BASELINE configs[1]: 1,000 tetrads, all-pairs list (494,500 tetrad pairs), random off — the default step and
    the brute-force distance kernels whose FP64 roofline SURVEY 8d defines."""
    import torch
    s = pkg.make_ring(1000, kind="ring", mole_cutoff=1e9)
    g = pkg.Engine(s, device=local)
    npairs = g.build_pairlist()
    crd0, vel0 = s.crd.copy(), s.vel.copy()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")   # > 126 MB L2
    stream = torch.cuda.ExternalStream(g.stream, device=torch.device("cuda", local))
    out = {"workload": "C1: synthetic 1000-tetrad replicated GC DNA ring, all-pairs NB (mole_Cutoff=1e9), random forces off",
           "tetrads": 1000, "tetrad_pairs": npairs, "l2": "256 MB device fill between timed steps"}
    flop = npairs * ATOMS * ATOMS * FLOP_PER_ATOM_PAIR
    for key, mode in (("default_step", 4), ("full_scan", 1), ("fp32_prefilter", 2)):
        g.set_nb_scan_mode(mode)
        g.set_state(crd0, vel0)
        g.step(3, 0)
        k = 20 if mode == 4 else 5
        tot = 0.0
        for i in range(k):
            g.mark(2 * i); g.step(1, 0); g.mark(2 * i + 1)
            with torch.cuda.stream(stream):
                flush.fill_(1)
        g.sync()
        ms = sum(g.mark_elapsed(2 * i, 2 * i + 1) for i in range(k)) / k
        g.set_state(crd0, vel0)
        ph = phase_times(g, lambda r: g.step(r, 0), 3)
        ent = {"nb_scan_mode": mode, "ms_per_step": ms, "value": npairs / (ms * 1e-3), "unit": UNIT,
               "nb_distance_ms": ph["nb_distance"]}
        if mode == 1:
            ach = flop / (ph["nb_distance"] * 1e-3) / 1e12
            ent.update({"bound": "fp64", "algorithmic_tflops": ach, "peak_tflops": peak_tf, "frac": ach / peak_tf,
                        "hw_flop_frac": ach / peak_tf * 6.0 / 9.0,
                        "note": "every one of the 63,504 atom pairs of every listed pair evaluated, like the reference; "
                                "9 algorithmic flops per atom pair (SURVEY 8d) in 3 DFMA = 6 hardware flops"})
        if mode == 2:
            ent.update({"bound": "fp32 (packed FFMA2 prefilter; the forces still come from the exact FP64 kernel)",
                        "atom_pairs_per_s": npairs * ATOMS * ATOMS / (ph["nb_distance"] * 1e-3),
                        "note": "BASELINE configs[4]'s FP32 variant of the distance pass: same selected pairs, bit-identical results"})
        out[key] = ent
    g.close()
    return out


def extra_contact(pkg, local, peak_tf):
    """Contact-bearing workload: a two-lap coil of 10,000 tetrads whose laps touch over ~15 % of the chain, against
    the clash-free ring of the same size; then the same with electrostatics on (qfac = 332.064/4, edmd.cpp:240),
    where every atom pair inside atom_Cutoff goes through the exact force kernel."""
    out = {"workload": "coil: 10000-tetrad two-lap closed coil (make_ring kind=coil, laps=2, lap_gap=40 A), cutoff pair list, random off"}
    for name, kind, qfac in (("ring", "ring", 0.0), ("coil", "coil", 0.0), ("ring_qfac", "ring", 332.064 / 4.0),
                             ("coil_qfac", "coil", 332.064 / 4.0)):
        s = pkg.make_ring(10000, kind=kind, laps=2, lap_gap=40.0)
        g = pkg.Engine(s, device=local)
        g.set_nb_consts(100.0, qfac)
        npairs = g.build_pairlist()
        crd0, vel0 = s.crd.copy(), s.vel.copy()
        g.step(3, 0); g.sync()
        g.set_state(crd0, vel0)
        ms = timed_steps(g, lambda r: g.step(r, 0), 20, 0)
        st = g.nb_stats()
        g.set_state(crd0, vel0)
        ph = phase_times(g, lambda r: g.step(r, 0), 5)
        ent = {"tetrad_pairs": npairs, "qfac": qfac, "ms_per_step": ms, "kernel_ms": ph, **st,
               "marked_fraction": st["marked_tetrads"] / s.n}
        # both owners evaluate a flagged pair's masked blocks: 2 x 1024 atom pairs per block
        evals = 2 * 1024 * st["force_blocks"]
        if evals:
            ent["force_pass_atom_pair_evals"] = evals
        out[name] = ent
        g.close()
    out["coil_over_ring"] = out["coil"]["ms_per_step"] / out["ring"]["ms_per_step"]
    # the exact force kernel alone (qfac != 0): finish phase minus the streaming integrate of the same size
    fin = out["coil_qfac"]["kernel_ms"]["finish"] - out["ring"]["kernel_ms"]["finish"]
    ev = out["coil_qfac"].get("force_pass_atom_pair_evals", 0)
    if fin > 0 and ev:
        out["force_kernel_qfac"] = {"ms": fin, "atom_pair_evals_per_s": ev / (fin * 1e-3),
                                    "note": "exact expression tree (no FMA, IEEE division), ~10 FP64 instructions per "
                                            "evaluated atom pair + ~25 per pair inside atom_Cutoff; FP64-pipe activity from ncu "
                                            "is in profiles/"}
    return out


def extra_c2(pkg, local):
    """BASELINE configs[2]: 10,000-tetrad ring, shipped cutoffs, Langevin random forces on (1 GPU here)."""
    s = pkg.make_ring(10000, kind="ring")
    g = pkg.Engine(s, device=local)
    g.set_rng(RNG_TIME0, 1, 0)
    npairs = g.build_pairlist()
    crd0, vel0 = s.crd.copy(), s.vel.copy()
    g.step(5, 1); g.sync()
    g.set_state(crd0, vel0)
    ms = timed_steps(g, lambda r: g.step(r, 1), 100, 0)
    g.set_state(crd0, vel0)
    ph = phase_times(g, lambda r: g.step(r, 1), 5)
    g.close()
    return {"workload": "C2: synthetic 10000-tetrad replicated GC DNA ring, cutoff pair list, Langevin random forces on, 1 GPU",
            "tetrads": s.n, "tetrad_pairs": npairs, "ms_per_step": ms, "value": npairs / (ms * 1e-3), "unit": UNIT,
            "steps_per_sec": 1e3 / ms, "kernel_ms": ph}


def multi_gpu_parity(pkg, dmod, dist, torch, rank, world, local):
    """Sharded engine vs one GPU, bit for bit, on a small contact-bearing system (outside the timed region)."""
    import numpy as np
    s = pkg.make_ring(200, kind="coil", laps=4, lap_gap=9.0)
    g = pkg.Engine(s, device=local)
    sh = dmod.ShardedEngine(dmod.EngineBackend(g, torch.device("cuda", local)), dist, s.n, rank, world)
    sh.set_rng(RNG_TIME0, 1, 0)
    npairs = sh.build_pairlist()
    sh.step(5, 1); sh.merge(); sh.build_pairlist(); sh.step(3, 1); sh.finish()
    c, v = g.get_state()
    ok = True
    if rank == 0:
        g1 = pkg.Engine(s, device=local)
        g1.set_rng(RNG_TIME0, 1, 0)
        ok = g1.build_pairlist() == npairs
        g1.step(5, 1); g1.merge(); g1.build_pairlist(); g1.step(3, 1)
        c1, v1 = g1.get_state()
        ok = ok and bool(np.array_equal(c, c1) and np.array_equal(v, v1))
        g1.close()
    sh.close(); g.close()
    return ok


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-sample", type=int, default=20000, help="tetrads in the CPU baseline's sample of C3")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the C1 / contact-coil measurements beside the headline")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    pkg = load_pkg()
    eng_mod = importlib.import_module("msc-project_b200.engine")
    dist_mod = importlib.import_module("msc-project_b200.dist")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the EDMD engine has no CPU fallback")
    torch.cuda.set_device(local)
    numa = bind_near_gpu(local) if world > 1 else None      # (one rank: keep every host thread for the CPU baseline)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    warmup = max(args.warmup, 3)
    steps = max(args.steps, 1)
    s = pkg.make_ring(N_C3, kind="ring")
    g = pkg.Engine(s, device=local)
    sh = dist_mod.ShardedEngine(dist_mod.EngineBackend(g, torch.device("cuda", local)), dist, s.n, rank, world)
    sh.set_rng(RNG_TIME0, 1, 0); g.set_rng(RNG_TIME0, 1, 0)
    npairs = sh.build_pairlist()
    assert npairs == ring_pair_count(s), (npairs, ring_pair_count(s))
    peak_tf, _ = eng_mod.measure_fp64_peak(local)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(k):
        if world == 1:
            g.step(k, 1)             # edmd_step: the public entry point (replays the captured step graph)
        else:
            sh.step(k, 1)            # the same call on every rank: the sharded step lives inside the engine

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # started early: nvidia-smi needs ~100 ms before its first sample

    run_steps(warmup)
    g.sync()

    # ---- timed region: K steps back to back, device time by CUDA events on the engine stream, max over ranks.
    # The state (3.1 GB of planes per GPU) is far larger than the 126 MB L2, so nothing is flushed between steps.
    launches0 = g.launch_count
    barrier()
    sampler.mark_begin()
    g.mark(0); run_steps(steps); g.mark(1)
    barrier()
    total_ms = max_over_ranks(g.mark_elapsed(0, 1))
    sampler.mark_end()
    launches = g.launch_count - launches0
    ms_per_step = total_ms / steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-phase breakdown (separate short pass: the timers put event records between the kernels)
    ph = phase_times(g, lambda r: run_steps(r), 5)
    barrier()
    st = g.nb_stats()

    # ---- the reference's outer loop (simulation.cpp:34-55): ntsync = 100 steps, merge, pair-list rebuild; the second of
    # two blocks is timed (the first pays one-off costs: lazily loaded kernels, first peer copies, graph re-instantiation)
    block_ms = rebuild_ms = None
    for _ in range(2):
        barrier()
        t0 = time.perf_counter()
        run_steps(100); g.sync()
        barrier()
        t1 = time.perf_counter()
        sh.merge(); sh.build_pairlist(); g.sync()
        barrier()
        t2 = time.perf_counter()
        block_ms = max_over_ranks((t2 - t0) * 1e3)
        rebuild_ms = max_over_ranks((t2 - t1) * 1e3)

    # ---- e2e: through the C ABI with HOST buffers, one step per call.  One GPU: edmd_step_host on the caller's
    # pinned Tetrad arrays (upload -> step -> download).  Several GPUs: every rank moves its OWN block of tetrads
    # (edmd_set_state_owned / edmd_get_state_owned) around the same sharded step.
    ke = max(3, min(steps, 10))
    nbytes_all = 2 * s.crd.size * 8
    if world == 1:
        for _ in range(2):
            g.step_host(1, 1)
        t0 = time.perf_counter()
        for _ in range(ke):
            g.step_host(1, 1)
        t_e2e = (time.perf_counter() - t0) / ke
    else:
        sh.finish()
        g.get_state(copy=False)
        for _ in range(2):
            g.push_state_owned(); sh.step(1, 1); g.get_state_owned()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ke):
            g.push_state_owned(); sh.step(1, 1); g.get_state_owned()
        barrier()
        t_e2e = max_over_ranks((time.perf_counter() - t0) / ke)
    e2e = {"value": npairs / t_e2e, "unit": UNIT, "h2d_bytes_per_step": nbytes_all, "d2h_bytes_per_step": nbytes_all,
           "ms_per_step": t_e2e * 1e3,
           "how": ("edmd_step_host: host Tetrad arrays (pinned) -> device, one step, device -> host Tetrad arrays" if world == 1 else
                   "per rank: edmd_set_state_owned (its block of the host arrays) -> sharded edmd_step -> edmd_get_state_owned; "
                   "bytes are the total over ranks"),
           "host_cpus_rank0": numa}

    parity = None
    if world > 1:
        parity = multi_gpu_parity(pkg, dist_mod, dist, torch, rank, world, local)
    owned = sh.plan.t1 - sh.plan.t0 if sh.plan is not None else s.n
    sh.close()
    g.close()

    extras = {}
    if world == 1 and not args.no_extras:
        extras["c1"] = extra_c1(pkg, eng_mod, local, peak_tf)
        extras["contact"] = extra_contact(pkg, local, peak_tf)
        try:
            extras["c2"] = extra_c2(pkg, local)
        except Exception as ex:          # an extra must never cost the bench line
            extras["c2"] = {"error": repr(ex)}

    if rank == 0:
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); hbm_src = "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            hbm_peak = 6541.1; hbm_src = "fallback: the value MEASURED_PEAKS.json held when this was written"
        # per-kernel roofline of the timed step (rank 0's share of the tetrads / pairs)
        kernels = {}
        for nm in ("superpose", "project", "random", "finish"):
            if ph.get(nm):
                gbs = owned * BYTES[nm] / (ph[nm] * 1e-3) / 1e9
                kernels[nm] = {"ms": ph[nm], "bound": "hbm", "bytes_per_tetrad": BYTES[nm], "achieved": gbs, "peak": hbm_peak,
                               "unit": "GB/s", "frac": gbs / hbm_peak}
        if ph.get("nb_distance"):
            # the culled distance pass evaluates only the 32 x 32 atom blocks the bounding spheres cannot exclude
            ev = st.get("sphere_blocks", 0) * 1024
            tf = ev * FLOP_PER_ATOM_PAIR / (ph["nb_distance"] * 1e-3) / 1e12
            kernels["nb_distance"] = {"ms": ph["nb_distance"], "bound": "fp64", "evaluated_atom_pairs": ev,
                                      "achieved": tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": tf / peak_tf,
                                      "note": "three launches (bounding spheres, pair selection, masked distance kernel); flops "
                                              "counted for EVALUATED atom pairs only, 9 per pair"}
        dom = max(kernels, key=lambda k: kernels[k]["ms"]) if kernels else None
        roof = {"kernel": dom, **{k: v for k, v in (kernels.get(dom) or {}).items() if k != "ms"},
                "ms_per_launch": (kernels.get(dom) or {}).get("ms"), "traffic": None,
                "traffic_note": "not measured inside this run; ncu dram bytes per launch are in profiles/",
                "peak_source": hbm_src + "; FP64: DFMA microbenchmark edmd_measure_fp64_peak, measured in this run",
                "kernels": kernels}
        if "c1" in extras and "full_scan" in extras["c1"]:
            fs = extras["c1"]["full_scan"]
            roof["nb_full_scan"] = {"kernel": "nb_scan_kernel<1> on C1 (every atom pair of every listed pair, like the reference)",
                                    "bound": "fp64", "achieved": fs["algorithmic_tflops"], "peak": peak_tf, "unit": "TFLOP/s",
                                    "frac": fs["frac"], "hw_flop_frac": fs["hw_flop_frac"], "ms_per_launch": fs["nb_distance_ms"]}
        line = {
            "metric": METRIC, "value": npairs / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "steps_per_sec": 1e3 / ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(npairs),
            "parallelism": f"{world} GPU(s): contiguous tetrad blocks x equal pair slices; exchanges over NVLink peer memory "
                           "inside the step graph" if world > 1 else "1 GPU",
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "kernel_ms": ph,
            "nb_stats": st,
            "ntsync_block": {"steps": 100, "ms_total_incl_merge_and_pairlist_rebuild": block_ms,
                             "ms_merge_and_pairlist_rebuild": rebuild_ms, "ms_per_step": block_ms / 100.0},
            "roofline": roof,
        }
        if parity is not None:
            line["multi_gpu_parity"] = parity
        if extras:
            line["extras"] = extras
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_sample(s, npairs, args.cpu_sample, with_variants=True)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
