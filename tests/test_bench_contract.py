"""bench.py contract pieces that run without a GPU: the reference arm prints ONE JSON line with the
contract's keys (the reference's own CPU code timed through the oracle library), and the GPU arm
refuses to run without a device instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, env=None):
    e = dict(os.environ); e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, env=e)


def test_reference_arm_prints_the_contract_line():
    r = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "500")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "tetrad_pair_interactions_per_sec" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["unit"] == "tetrad-pairs/s" and d["steps"] == 1 and d["dtype"] == "f64"
    assert d["config"]["tetrad_pairs"] == 400005 and d["config"]["tetrads"] == 100000 and "C3" in d["config"]["workload"]
    sys.path.insert(0, ROOT)
    import importlib
    bench = importlib.import_module("bench")
    assert d["config"] == bench.config_dict(400005)          # the GPU arm prints the same dict
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_exit_quietly():
    r = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "500", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a CUDA device is present")
    r = run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_clock_sampler_parses_nvidia_smi_lines():
    """The `clocks` object: median SM clock and throttle reasons of the samples inside the timed region,
    falling back to every sample of the run when the region is shorter than a few sampling periods."""
    sys.path.insert(0, ROOT)
    import importlib, time, datetime
    bench = importlib.import_module("bench")
    cs = bench.ClockSampler(0)
    cs.proc = type("P", (), {"terminate": lambda s: None, "wait": lambda s, timeout=None: 0, "kill": lambda s: None})()
    now = time.time()

    def line(dt, mhz, power_cap="Not Active", hw="Not Active"):
        ts = datetime.datetime.fromtimestamp(now + dt).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]
        return f"{ts}, {mhz}, 1965, 600.0, 0x0, {hw}, Not Active, Not Active, {power_cap}"
    cs.lines = [line(-1.0, 1200), line(0.1, 1965), line(0.2, 1950, power_cap="Active"), line(0.3, 1965), line(2.0, 900, hw="Active"),
                "garbage line", "2026/01/01 00:00:00.000, [N/A], [N/A], x, x, x, x, x, x"]
    cs.t_begin, cs.t_end = now, now + 0.4
    c = cs.stop()
    assert c["window"] == "timed region" and c["samples"] == 3 and c["sm_mhz"] == 1965 and c["sm_max_mhz"] == 1965
    assert c["reasons"] == ["sw_power_cap"]
    cs.t_begin, cs.t_end = now + 0.15, now + 0.16            # a 10 ms region: only one sample inside
    c = cs.stop()
    assert c["window"].startswith("whole run") and c["samples"] == 5 and "hw_slowdown" in c["reasons"]


def test_ring_pair_count_equals_the_oracle_list():
    """bench.ring_pair_count (used by both arms for the workload's pair count) == the reference pair list."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import importlib
    import pyoracle
    bench = importlib.import_module("bench")
    pkg = importlib.import_module("msc-project_b200")
    for n in (90, 400, 1000):
        s = pkg.make_ring(n, kind="ring")
        o = pyoracle.Oracle(s, "port")
        assert o.build_pairs() == bench.ring_pair_count(s)


def test_cpu_baseline_variants():
    """The GPU arm's cpu_baseline also reports BASELINE.md section 3's other two CPU figures — one core, and the
    reference at the -O0 it ships with — each on a bounded sample; a failing variant yields an error entry, never
    an exception (it must not cost the bench line)."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import importlib
    import pyoracle
    bench = importlib.import_module("bench")
    pkg = importlib.import_module("msc-project_b200")
    s = pkg.make_ring(2000, kind="ring")
    npairs = bench.ring_pair_count(s)
    kind = "reference" if pyoracle.available("reference") else "port"
    r = bench.cpu_sample(s, npairs, 1600, kind_wanted=kind, reps=1, with_variants=True)
    assert r["value"] > 0 and r["kind"] == kind
    v = r["variants"]
    one = v["one_core"]
    assert one["cores"] == 1 and 0 < one["value"] <= r["value"] * 1.5
    o0 = v["shipped_O0"]
    if kind == "reference" and pyoracle.available("reference_O0"):
        assert o0["cores"] == r["cores"] and 0 < o0["value"] < r["value"]      # -O0 is slower than -O3
    else:
        assert "unavailable" in o0
    assert bench.cpu_sample(s, npairs, 1600, kind_wanted=kind, reps=1)["variants"] is None
