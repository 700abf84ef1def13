"""Pins bench.py's roofline accounting to the SURVEY.md §8d figures and checks the reference arm prints the contract line."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_algorithmic_bytes_match_survey():
    import bench
    from onebit_b200 import LLAMA2_13B, LLAMA_7B
    b7 = bench.bitlinear_bytes(LLAMA_7B, 1)
    b13 = bench.bitlinear_bytes(LLAMA2_13B, 1)
    assert abs(b7["packed_per_step"] / 1e6 - 809.5) < 0.1      # 809.5 MB of packed signs (7B)
    # §8d: 814.5 MB = packed + g/h; the per-call formula also counts fp16 x and y (another 5.0 MB at B = 1)
    assert abs(b7["per_step"] / 1e6 - (814.5 + 5.0)) < 0.1
    assert abs(b13["packed_per_step"] / 1e6 - 1586.0) < 0.5
    assert abs(b13["per_step"] / 1e6 - (1593.8 + 7.8)) < 0.1
    assert b7["gemv_launches"] == 128 and b13["gemv_launches"] == 160


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, cwd=str(ROOT), timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "tok/s" and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
    # both arms print the same `config` object for the same workload; the extrapolation is said out loud
    import bench
    assert line["config"] == bench.workload_config("LLaMA-7B-OneBit", 1, 1, False)
    assert "EXTRAPOLATED" in line["note"] and line["metric"] == bench.METRIC
