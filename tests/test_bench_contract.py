"""bench.py's reference arm runs without a GPU: check the one-JSON-line contract (keys, types, a clean stdout)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--batch", "20"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout                                   # stdout carries the JSON line and nothing else
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "arch1_train_samples_per_s" and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 0 and abs(d["ms_per_step"] * d["value"] / 1e3 - 20) < 1e-6 * 20      # 20 samples per step
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == d["value"] and cb["cores"] >= 1 and "sample" in cb
    assert "workload" in d["config"] and "model" not in d["config"]


def test_other_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--batch", "20"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_our_arm_fails_loudly_without_a_gpu():
    import novel_vqa_b200 as nv
    if nv.device_count() > 0:
        import pytest
        pytest.skip("a B200 is visible: the arm runs")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                       text=True, timeout=600, cwd=ROOT)
    assert p.returncode != 0 and p.stdout.strip() == "" and "no CPU fallback" in p.stderr


def test_reference_arm_uses_all_host_threads_under_torchrun_env():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the round-1 SCALE ratios at N >= 2 were taken against a
    single-threaded CPU arm.  The reference arm sets the torch thread count itself."""
    try:
        want = len(os.sched_getaffinity(0))
    except AttributeError:
        want = os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--batch", "20"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["cores"] == want and d["n_gpus"] == 2
