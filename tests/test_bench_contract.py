"""CPU: the reference arm of bench.py prints exactly one JSON line with the contract's keys (the GPU arm needs a
device; its line carries the same keys plus `roofline`, `clocks`, `gpu_launches`)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"stdout must hold exactly one JSON line, got {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_json_contract():
    line = run("--impl", "reference", "--workload", "cfg1", "--steps", "2", "--warmup", "1")
    assert line["steps"] == 2 and line["warmup"] == 1
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["value"] > 0 and line["higher_is_better"] is True


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--workload", "cfg1", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_falls_back_to_the_port_without_the_installed_reference(tmp_path, monkeypatch):
    """Without baseline/_ref the CPU arm is the oracle port and says so."""
    import importlib
    sys.path.insert(0, ROOT)
    bw = importlib.import_module("bench_workloads")
    monkeypatch.setattr(bw, "reference_package", lambda: None)
    wl = bw.WORKLOADS["cfg1"](0, 1, None)
    res = wl.reference_arm(steps=1, warmup=0)
    assert res["kind"] == "port" and res["value"] > 0 and res["ms_per_step"] > 0


def test_configs_do_not_depend_on_the_run():
    """`config` must be identical in the GPU arm and the reference arm: a pure function of the workload definition."""
    import importlib
    sys.path.insert(0, ROOT)
    bw = importlib.import_module("bench_workloads")
    for name, cls in bw.WORKLOADS.items():
        assert cls(0, 1, None).config() == cls(3, 8, None).config(), name
        assert "workload" in cls(0, 1, None).config()
    assert bw.DEFAULT_WORKLOAD == "eval_cfg5"
