"""The measurement contract of bench.py that can be checked without a GPU: the reference arm prints ONE JSON line with the keys the
driver reads, and its `config` equals the config this repository's arm prints for the same command line (the driver compares them)."""
import argparse
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def test_reference_arm_line_and_shared_config():
    res = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decode_tokens_per_s" and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] >= 1 and d["ms_per_step"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    import bench

    args = argparse.Namespace(batch=1, tp_collective="fused", replicas=False)
    mine = bench.decode_config(args, bench.workload_name("1b", "bf16", 1, 0), 1, False)
    assert d["config"] == mine
    assert "configs[1]" in mine["workload"] and mine["kv_len"] == 512 and "L2" in mine["l2"]
    # tensor parallel: both arms name the same sharded workload
    tp = bench.decode_config(args, bench.workload_name("8b", "bf16", 1, 0), 8, True)
    assert tp["parallelism"].startswith("tp8: ONE model") and "configs[3]" in tp["workload"]
