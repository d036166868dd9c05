"""CPU-side checks of bench.py: the reference arm runs without a GPU and prints the contract's JSON line."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("config", ["c2", "c3", "c4"])
def test_reference_arm_prints_contract_line(config):
    size = "20000" if config == "c3" else str(2 << 20)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", config,
                          "--steps", "1", "--warmup", "0", "--size", size], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "input GB/s scanned" and line["unit"] == "GB/s"
    assert line["value"] > 0 and line["higher_is_better"] is True and line["gpu_launches"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["config"] == config


def test_bench_refuses_to_run_the_product_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--size", "1048576"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode != 0          # no CPU fallback: the product arm fails loudly
