"""The real NCCL path on hardware: avirb200_resize_sharded under one process per GPU must give
the 1-GPU bits (needs >= 2 visible GPUs; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_nccl.py -m gpu`)."""
import os
import subprocess
import sys

import pytest

import avir_b200 as ab

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("nranks", [2, 4, 8])
def test_nccl_sharded_matches_single_gpu(nranks):
    if _gpus() < nranks:
        pytest.skip("needs %d GPUs" % nranks)
    assert ab.device_count() >= nranks
    port = 29500 + (os.getpid() % 200) + nranks
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nranks),
                        "--master-addr", "127.0.0.1", "--master-port", str(port),
                        os.path.join(ROOT, "tests", "nccl_worker.py")],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:]
    assert "mismatches=" in r.stdout
