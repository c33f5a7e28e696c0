"""Real multi-rank run (one process per GPU, NCCL): the sharded entry points of softgnss_python_b200.dist return the
bytes of the single-GPU run -- acquisition split by PRN, tracking and the navigation chain split by recording, result
gathers GPU to GPU.  Needs two visible GPUs (`gpurun --gpus 2`); the CPU-side gather logic is covered by
tests/test_dist_gloo.py."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_results_are_byte_identical_to_one_gpu(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`); %d visible" % n)
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multirank_worker.py")]
    env = dict(os.environ, SGX_MR_OUT=str(tmp_path))
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT, env=env)
    files = sorted(tmp_path.glob("rank*.json"))
    lines = [json.loads(f.read_text()) for f in files]
    assert res.returncode == 0 and len(lines) == world and all(x.get("ok") for x in lines), \
        (res.stdout[-3000:], res.stderr[-3000:])
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "multirank_test.json"), "w") as f:
            json.dump(lines, f)
