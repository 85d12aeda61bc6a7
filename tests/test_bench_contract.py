"""bench.py contract on the CPU: the reference arm prints ONE JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "solves/s" and d["higher_is_better"] is True and d["value"] > 0
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "workload" in d["config"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)


def test_committed_bench_lines_carry_the_contract_keys():
    """The bench lines committed under profiles/ (taken on B200 boxes) have every key of the contract: the driver-facing ones, `e2e` with byte counts,
    `roofline` with traffic, `cpu_baseline` with the host's effective parallelism, `clocks` without throttle reasons, and a non-zero launch count."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_*gpu_cfg[123]*.json")))
    assert files
    for f in files:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                  "e2e", "gpu_launches", "roofline", "clocks"):
            assert k in d, (f, k)
        assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64" and d["data"] == "synthetic" and d["warmup"] >= 3
        assert d["gpu_launches"] > 0 and "workload" in d["config"] and "model" not in d["config"]
        assert d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] <= d["value"]
        r = d["roofline"]
        assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["traffic"]
        assert d["clocks"]["reasons"] == [] and d["clocks"]["sm_mhz"] >= 0.9 * d["clocks"]["sm_max_mhz"]
        assert abs(d["value"] - d["config"]["instances_total"] * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3)) < 1e-6 * d["value"]
        if d["n_gpus"] == 1 and "cpu_baseline" in d:
            c = d["cpu_baseline"]
            assert c["kind"] == "port" and c["cores"] >= 1 and c["torque_rel_err_vs_gpu"] < 1e-4
            if f.endswith("_final.json"):            # lines from the start of the round predate the host-parallelism record
                assert c["host_parallelism"]["effective"] >= 1
        if d["n_gpus"] > 1:
            assert d["gather"]["finite"] is True and d["gather"]["rows_at_rank0"] == d["config"]["instances_total"]
