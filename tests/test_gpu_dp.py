"""Multi-GPU data parallelism on real GPUs (self-skips with fewer than two): batch-sharded
gradients + the all-reduce inside `yota_program_run` against the single-GPU full-batch gradients,
for ncclAllReduce and for the library's own NVSwitch kernel, including evals issued back to back
(reductions of eval i overlapping the sweep of eval i+1).  Worker: tests/dp_gpu_check.py."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import gpu_count

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_sharded_gradients_with_allreduce_match_the_full_batch():
    n = gpu_count()
    if n < 2:
        pytest.skip(f"needs >= 2 GPUs on one box (found {n})")
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "dp_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"DP_GPU_OK dest_sharded_scatter nranks={n}" in r.stdout, r.stdout[-3000:]
    for mode in ("fp32", "tf32"):
        assert f"DP_GPU_OK mode={mode} allreduce=nccl nranks={n}" in r.stdout, r.stdout[-3000:]
