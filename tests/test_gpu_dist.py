"""-m gpu, needs >= 2 GPUs: one process per GPU over NCCL (SyncBN + bucketed gradient all-reduce) must reproduce the
single-GPU run of the same global batch."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("nvl,model", [("1", "UNet"), ("0", "UNet"), ("1", "UNet_attention")])
def test_data_parallel_matches_single_gpu_global_batch(nvl, model):
    """nvl=1: SyncBN statistics through the NVLink peer-memory kernel; nvl=0: through NCCL. UNet_attention: the gates' three
    BatchNorms each (statistics rows of the GEMM epilogues, fp64 sums of the gate kernels) under SyncBN."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_worker.py")]
    # nvl=1 also replays the data-parallel step as CUDA graphs (opt-in, B200UNET_DP_GRAPHS=1) against the eager launches
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300,
                       env=dict(os.environ, B200UNET_NVL_SYNCBN=nvl, B200UNET_DP_GRAPHS="1", DP_MODEL=model))
    sys.stdout.write(r.stdout[-4000:])
    sys.stderr.write(r.stderr[-4000:])
    assert r.returncode == 0 and "DP_OK" in r.stdout
