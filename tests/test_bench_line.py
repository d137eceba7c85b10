"""bench.py (our arm) at a small size on the GPU box: ONE JSON line with every key of the bench contract."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def test_bench_line_has_the_contract_keys():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--log2n", "22", "--steps", "3", "--warmup", "3",
                          "--cpu-baseline-log2n", "22"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["dtype"] == "u32" and d["value"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and 0 < r["frac"] < 1.5 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert "traffic" in r
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] == 4 << 22 and e["d2h_bytes_per_step"] == 4 << 22
    assert e["pipelined"]["value"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    assert d["gpu_launches"] > 0
    assert d["parity"]["bit_exact_vs_reference_cpu_sort"] is True and d["parity"]["e2e_bit_exact_vs_reference_cpu_sort"] is True
    assert set(("sm_mhz", "sm_max_mhz", "reasons")) <= set(d["clocks"])
