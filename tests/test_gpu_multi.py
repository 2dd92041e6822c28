"""ONE raster cut along its drainage graph over 2 processes must reproduce the single-GPU result bit for bit (the
reference's own decomposability evidence: tests/test_subcatchments.py:111-112).  Runs tools/run_dist_check.py under
torch.distributed.run: with two GPUs the ranks exchange over NVLink peer memory (and NCCL plumbing); on a one-GPU box
both ranks share the device (CUDA IPC on the same device, gloo plumbing) -- the same kernels and exchange protocol."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode", ["parity", "stress"])
def test_ldd_cut_bit_identical_two_ranks(gpu_lib, mode):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "run_dist_check.py")]
    if mode == "stress":
        cmd.append("--stress")
    env = dict(os.environ, NCCL_DEBUG="WARN")
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-6000:])
    assert "DIST CHECK PASSED" in r.stdout, r.stdout[-3000:]
