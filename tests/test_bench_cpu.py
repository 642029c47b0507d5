"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line on stdout (everything
else, including what libraries write to fd 1, goes to stderr) with the keys the driver reads."""

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ, OMP_NUM_THREADS="4", DFM_BENCH_CPU_SMALL="1")     # bounded sample: keeps the CPU suite short
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("train samples/sec") and d["unit"] == "samples/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["workload"].startswith("deepfm_criteo")
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and "batch" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_fp32_emulation_switch_refuses_after_torch_import():
    import torch  # noqa: F401
    from deepfm_b200 import fp32_emulation as E
    if not E._state["enabled"]:
        assert E.enable() is False and "torch was imported first" in E.status()
