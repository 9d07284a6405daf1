"""bench.py contract pieces that can be checked without a GPU: the reference arm (CPU port of the reference's
literal algorithm) prints exactly one JSON line with the required keys."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ntimes", "100", "--batch", "64", "--cpu-pulses-per-thread", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "evals/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["value"] > 0 and d["metric"].startswith("GRAPE cost+grad evals/sec")


def test_flop_and_degree_helpers():
    sys.path.insert(0, str(ROOT))
    import bench
    assert bench.canonical_flops(5, 1000, 1, 1, 0) == 1.4e7          # 1.40e7 per C4 evaluation
    assert abs(bench.canonical_flops(5, 1000, 1, 1, 1) - 4.2e7) < 1
    assert bench.taylor_degree_for(5.4e-3) == 5 and bench.taylor_degree_for(0.05) == 8
    assert bench.k_steps_flops(1000, 5, False) == 5 * 4 * 75 * 8 * 1000
