"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`) prints one JSON
line with the keys the driver reads, on the headline workload; without a CUDA device the B200 arm
refuses to run (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=600, cwd=ROOT)


def test_reference_arm_json_line():
    r = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--skip-extra")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1  # ONE JSON line
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "aggregated_edges_per_sec" and d["unit"] == "edges/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1
    assert d["config"]["workload"].startswith("sage_reddit") and d["config"]["edges_per_step"] == 281600
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "minibatch" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same `config` object as the B200 arm prints (the driver compares them key by key)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.sage_config(1)


def test_reference_arm_uses_all_host_threads_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1 to its ranks; the CPU arm must still use every core, honour
    --steps/--warmup and report them (the N>=2 reference lines ran single-threaded)."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
                        "--warmup", "3", "--skip-extra"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() or 1)
    assert d["steps"] == 2 and d["warmup"] == 3 and d["n_gpus"] == 2


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_b200_arm_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        return  # on the GPU box the arm runs for real (bench.py itself)
    r = _run("--steps", "1", "--warmup", "1", "--skip-extra", "--no-cpu-baseline")
    assert r.returncode != 0
    assert "no CPU fallback" in r.stdout
