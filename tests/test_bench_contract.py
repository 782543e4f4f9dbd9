"""CPU: the reference arm of bench.py (the CPU restatement timed on the host cores) prints one JSON line with the keys the
driver reads.  The GPU arm has the same keys plus `roofline` / `clocks`; it needs a B200 and is exercised by the driver."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                          "--envs", "512"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["value"] > 0 and d["steps"] == 3
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_other_ranks_of_the_reference_arm_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


import pytest  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("extra", [[], ["--log-wrapper"], ["--policy"], ["--streams", "2"]])
def test_gpu_arm_prints_the_contract_line(extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--envs", "2048", "--steps", "64", "--warmup", "3",
                          "--cpu-seconds", "0.5", "--e2e-steps", "64", "--min-ms", "0", "--stats-steps", "24", "--rollout-steps", "16"] + extra, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["value"] > 0 and d["gpu_launches"] > 0 and d["n_gpus"] == 1 and d["dtype"] == "u8"
    assert "workload" in d["config"] and set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    # one replayed graph = 2 x 16 steps (two oc_rollout launches; --policy / --streams use the one-launch-per-step graph: 2 x 8)
    block = 32 if (not extra or extra == ["--log-wrapper"]) else 16
    assert d["steps"] == 64 and d["replays"] == 64 // block and d["requested_steps"] == 64
    if "--policy" not in extra:
        w = d["stats_window"]
        assert w["stats_allreduced"][6] == 24 * 2048 and w["allreduce_equals_sum_of_shards"] and w["steps"] == 24
    if not extra:
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert r["kernel"] == "oc_env_kernel<MODE_ROLLOUT>" and r["steps_per_launch"] == 16 and 0 < r["frac"] < r["frac_survey_bytes"]
        assert abs(r["algorithmic_bytes_per_env_step"] - (52 * 20 + 21 + (6 * 20 + 82) / 16)) < 1e-9 and r["survey_8d_bytes_per_env_step"] == 1277
        e = d["e2e"]
        assert d["cpu_baseline"]["kind"] == "port" and e["wire"] == "packed" and e["h2d_bytes_per_step"] == 2048 and e["d2h_bytes_per_step"] == 2048
        assert e["value"] > 0 and e["value"] != d["value"] and e["value_reference_wire"] > 0


@pytest.mark.gpu
def test_gpu_arm_time_floor():
    """Whatever --steps says, the device-timed window is at least --min-ms long and the end-to-end leg at least 2,000 steps."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--envs", "4096", "--steps", "20", "--warmup", "3",
                          "--no-cpu", "--min-ms", "60", "--stats-steps", "8", "--rollout-steps", "16"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert d["timed_window_ms"] >= 55.0 and d["steps"] == 32 * d["replays"] and d["steps"] > 20
    assert abs(d["ms_per_step"] * d["steps"] - d["timed_window_ms"]) < 1e-6 * d["timed_window_ms"] + 1e-9
    assert d["e2e"]["steps"] >= 2000


def test_reference_arm_prefers_a_real_jax(monkeypatch, capsys):
    """With an importable jax + reference tree the arm times the reference's own jitted, vmapped step (kind
    "reference-jax-cpu"); the timing harness is stubbed here -- tests/test_real_jax.py runs the real one on the shim."""
    import types
    sys.path.insert(0, ROOT)
    import bench
    from baseline import ref_jax
    calls = []

    def fake_time(jax, ns, lay, name, *, n_envs, n_steps, **kw):
        calls.append((n_envs, n_steps))
        return 1.0e6 * n_envs / 1024, n_steps * 1e-3, 0.5
    monkeypatch.setattr(ref_jax, "available", lambda: True)
    monkeypatch.setattr(ref_jax, "import_jax", lambda: types.SimpleNamespace(__version__="0.0-test"))
    monkeypatch.setattr(ref_jax, "load_reference", lambda: None)
    monkeypatch.setattr(ref_jax, "time_random_rollout", fake_time)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--steps", "7", "--envs", "4096"])
    bench.run_reference(bench.parse())
    d = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference-jax-cpu" and d["steps"] == 7
    assert calls[0] == (1024, 8) and calls[1][1] == 7 and 256 <= calls[1][0] <= 4096
    assert d["e2e"]["value"] == d["value"] == d["cpu_baseline"]["value"]
