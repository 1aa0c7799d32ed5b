"""GPU (needs >= 2 devices, skipped otherwise): the row-sharded handle reproduces the single-GPU solve
(tools/check_sharded.py under torchrun, NCCL).  "er": no locality -> staged peer-copy exchange + column passes;
"torus": locality -> the product gathers the few remote rows directly from the owner's memory."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("graph,port", [("er", "29533"), ("torus", "29534")])
def test_sharded_matches_single_gpu(graph, port):
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", port, os.path.join(ROOT, "tools", "check_sharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, CHK_GRAPH=graph))
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and lines, r.stdout[-2000:] + r.stderr[-2000:]
    assert json.loads(lines[-1])["sharded_check"] == "ok"
