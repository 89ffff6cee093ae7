"""Multi-GPU data plane on hardware (needs >= 2 visible GPUs; skipped on a 1-GPU box): the fused all-gather
(NVLink peer stores from the QP kernel + arrival flags + credits) must equal an NCCL all-gather bit for bit."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_fused_gather_equals_nccl_all_gather_world2():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tools", "gather_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, EMP_CHECK_BATCH="1024"))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert '"all_ranks_verified": true' in r.stdout
