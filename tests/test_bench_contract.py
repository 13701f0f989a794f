"""bench.py's contract, as far as it can be checked without a GPU: the reference arm (the C port of the reference's MALA
path on the host cores) runs and prints ONE JSON line with the keys the driver reads; the two arms share their
`config`; the GPU arm refuses to run without CUDA instead of falling back to anything."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT, env=e)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "0", env={"OMP_NUM_THREADS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "chain-steps/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["e2e"] == {"value": d["value"], "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and "sample" in cb
    # torchrun exports OMP_NUM_THREADS=1: the arm must still use every host core (the N >= 2 ratios were void)
    assert cb["cores"] == (os.cpu_count() or 1)
    import bench
    assert d["config"] == bench.make_config(1) and d["metric"] == bench.METRIC
    assert d["config"]["local_steps_per_bench_step"] == 1000 and d["config"]["n_chains_per_gpu"] == 8192


def test_reference_arm_is_rank_zero_only():
    r = _run("--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
             env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("needs a machine without CUDA")
    r = _run("--steps", "1", "--warmup", "0")
    assert r.returncode != 0
    assert "{" not in r.stdout            # no metric line from any fallback
