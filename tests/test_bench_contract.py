"""bench.py's reference arm runs without a GPU (and without loading libfastq_gpu.so): its JSON line carries the keys the driver reads."""
import json
import os
import subprocess
import sys

import pytest

from _util import ROOT


@pytest.mark.parametrize("workload", ["illumina_pe", "illumina_se"])
def test_reference_arm_line(workload):
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "fastq_info")):
        pytest.skip("oracle/_ref/fastq_info not built")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-reads", "20480",
                        "--workload", workload], capture_output=True, text=True, env=env, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "fastq_info_validated_GBps" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 1 and d["n_gpus"] == 1 and d["scaling"] == "weak" and d["dtype"] == "u8" and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_does_not_load_the_product_library():
    src = open(os.path.join(ROOT, "bench.py")).read()
    arm = src[src.index('if a.impl == "reference":'):src.index("# ------------------------------------------------------------------ our arm")]
    assert "fastq_utils_b200 as fq" not in arm and ".lib()" not in arm and "import torch" in arm  # (torch only for the long-read layout)
