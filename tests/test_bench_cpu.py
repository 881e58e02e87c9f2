"""bench.py's CPU legs: the `--impl reference` arm prints exactly one JSON line with the contract's keys on rank 0 and nothing
on the other ranks (the driver launches it under torchrun for N > 1).  No GPU, no compute on the product path."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env):
    env = dict(os.environ, **extra_env)
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0",
           "--ref-blocks", "1", "--no-decoder"]
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)


def test_reference_arm_line():
    r = _run({"RANK": "0", "WORLD_SIZE": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    assert line["metric"] == "denoise_steps_per_sec" and line["unit"] == "steps/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["value"] > 0 and abs(line["ms_per_step"] * line["value"] - 1e3) < 1e-3 * 1e3
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_are_silent():
    r = _run({"RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.strip() == ""


def test_committed_bench_lines_carry_the_contract_keys():
    """the bench lines committed as evidence (profiles/r2_bench_n1_final.json at N = 1, r2_bench_n2_final.json under torchrun at N = 2) hold
    every key the bench contract names: metric / value / unit, n_gpus, ms_per_step, scaling, dtype, data, config.workload, the roofline and
    cpu_baseline objects, e2e with its byte counts, clocks with throttle reasons, gpu_launches -- and the denoise value is consistent with
    ms_per_step"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for name, n in (("r2_bench_n1_final.json", 1), ("r2_bench_n2_final.json", 2)):
        txt = open(os.path.join(root, "profiles", name)).read().strip().splitlines()[-1]
        d = json.loads(txt)
        assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
        assert d["metric"] and d["unit"] and d["dtype"] and "workload" in d["config"] and "model" not in d["config"]
        assert d["warmup"] >= 3 and d["steps"] >= 1 and d["vs_baseline"] is None
        assert abs(d["value"] - n * 1000.0 / d["ms_per_step"]) / d["value"] < 1e-3       # one prompt per GPU: steps/s = N / step time
        r = d["roofline"]
        assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s") and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-6
        assert set(r["attention"]) == {"self", "cross"} and "CUDA graph" in r["attention"]["self"]["timing"]
        e = d["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
        c = d["clocks"]
        assert c["sm_mhz"] > 0 and c["sm_max_mhz"] >= c["sm_mhz"] and isinstance(c["reasons"], list)
        assert not {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(c["reasons"])
        assert d["gpu_launches"] > 0
        if n == 1:
            b = d["cpu_baseline"]
            assert b["kind"] in ("reference", "port") and b["cores"] >= 1 and b["value"] > 0 and b["sample"]
        g = d["gaussians"]
        assert g["decoder_cuda_graph"] is True and g["decoder_launches_per_forward"] > 0 and g["e2e_gaussians_per_sec"] > 0
