"""bench.py's output contract, as far as it can be checked without a GPU: the reference arm (the oracle's
reference-literal algorithm on the host cores) prints ONE JSON line with the keys the driver reads, and the clock
sampler parses nvidia-smi rows.  CPU only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "windowed LS spectra/sec" and d["unit"] == "windows/s"
    assert d["value"] > 0 and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["config"]["workload"] == "cfg2_windowpsd" and d["config"]["windows_per_gpu"] == 2047
    sys.path.insert(0, ROOT)
    import bench

    assert d["config"] == bench.cfg2_config(4096)  # the GPU arm prints the same dict (the driver compares the two)
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["vs_baseline"] is None and d["dtype"] == "f64"


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env,
                         cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_clock_sampler_parses_nvidia_smi_rows():
    sys.path.insert(0, ROOT)
    import bench

    s = bench.ClockSampler(0)
    s.rows = ["210, 1965, 180.1, Not Active, Not Active, Not Active, Not Active",
              "1965, 1965, 900.2, Not Active, Not Active, Not Active, Active",
              "1950, 1965, 950.0, Not Active, Not Active, Not Active, Not Active",
              "garbage"]
    c = s.stop()
    assert c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"] and c["samples"] == 3
    assert c["sm_mhz"] >= 1950.0  # median of the under-load half
    assert bench.ClockSampler(0).stop() == {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}


def test_workload_generators_have_the_baseline_shapes():
    sys.path.insert(0, ROOT)
    import bench
    import numpy as np

    t, y, f, n = bench.make_cfg2(nsamp=1 << 14, nw=4, nf=128)
    assert n == 4096 and len(t) == len(y) == 1 << 14 and len(f) == 128 and f[0] == 0 and np.all(np.diff(t) >= 0)
    t, y, f = bench.make_cfg3()
    assert len(f) == 8192 and f[0] == 0 and len(t) == 16384
    t, y, f = bench.make_cfg1()
    assert len(t) == 4096 and len(f) == 2048 and f[0] == 0
    t, y, u, f, n = bench.make_cfg5(nsamp=1 << 14)
    assert n == 4096 and len(f) == 512 and len(t) == len(y) == len(u) == 1 << 14
