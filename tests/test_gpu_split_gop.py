"""BASELINE.json config 4's forced case on real GPUs: a GOP that continues on another GPU after an NCCL
broadcast of the Decoder (tools/split_gop_check.py).  Needs >= 2 GPUs; the single-GPU box skips it (the
host side of the exchange is covered by tests/test_multigpu_gloo.py)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_gop_continues_bit_exactly_on_another_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(ROOT, "tools", "split_gop_check.py")], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-3000:]
    row = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert row["mismatches"] == 0 and row["backend"] == "nccl"
