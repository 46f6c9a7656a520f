"""bench.py pieces that run without a GPU: the reference arm (the CPU oracle port timed on the host
cores) prints one JSON line with the contract's keys, and the algorithmic byte count matches SURVEY.md 8d."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_per_unit():
    sys.path.insert(0, ROOT)
    import bench
    total, per = bench.alg_bytes_per_unit(bench.CONFIGS["C2"])
    assert abs(total - (2 * (4 * 64 + 16) + 32 / 11.0)) < 1e-9  # 547 B/unit, two-grid path, rho = 4
    assert per["pass1"] == per["pass2"] == 256.0
    total1, _ = bench.alg_bytes_per_unit(bench.CONFIGS["C1"])  # 342 x 341 padded to 1024^2: rho = 8.99
    assert 1180 < total1 < 1190


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS=str(min(8, os.cpu_count() or 1)))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "C1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "units/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["gpu_launches"] == 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
