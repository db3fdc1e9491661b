"""Sharded (one process per GPU) solve against the reference's record; needs >= 2 GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_two_rank_partial_schur(gpu):
    ndev = gpu.ab200_device_count()
    if ndev < 2:
        # explicit opt-out, not a silent pass: on a 1-GPU box the sharded parity is carried by
        # bench.py's `parity` block instead (it runs the golden solves at the rank count of
        # every scaling run and writes R, R_ref, max_rel into the JSON line the driver records)
        pytest.skip("needs >= 2 GPUs on the box (gpurun --gpus 2); sharded parity is also "
                    "recorded by `bench.py --gpus N` in its `parity` block")
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517",
           os.path.join(here, "mgpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    sys.stdout.write(out.stdout[-4000:])
    sys.stderr.write(out.stderr[-4000:])
    assert out.returncode == 0 and "[mgpu] OK" in out.stdout
