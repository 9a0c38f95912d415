"""N-rank gradient parity on real GPUs (needs >= 2 devices; the single-GPU driver box skips it): the all-reduced gradient on N
ranks equals the single-process gradient of the same per-rank loss — and, with global_dice, of the concatenated batch as the
reference's nn.DataParallel computes it (trainer.py:37-38, :55-57).  Runs tools/ddp_grad_check.py under torch.distributed.run."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("extra", [["--dtype", "fp32"], ["--dtype", "bf16"], ["--dtype", "fp32", "--global-dice"]])
def test_two_rank_allreduced_gradient_equals_single_process(extra):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with `gpurun --gpus 2`; log committed under profiles/)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "ddp_grad_check.py")] + extra
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=ROOT)
    print(p.stdout[-3000:])
    assert p.returncode == 0, p.stdout[-3000:]
    assert "-> OK" in p.stdout
