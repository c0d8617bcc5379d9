"""Multi-GPU tests (skipped with fewer than two devices): the row-sharded self-convection, the sweep split and the
flow-field slabs on real GPUs under torchrun / NCCL, checked against single-GPU evaluation and the oracle.  The
host-side partition logic is covered on CPU by tests/test_sharded_gloo.py."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _torchrun(n, script_args, timeout=900):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", str(port)] + script_args
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout)


@pytest.fixture(scope="module")
def ngpus():
    n = _ngpus()
    if n < 2:
        pytest.skip("needs at least two CUDA devices")
    return 2 if n < 4 else 4


def test_sharded_self_convection_bitwise_equal_to_one_rank(ngpus):
    """scripts/dist_parity.py: G-rank self-convection (fused peer-store all-gather, NCCL all-gather; fast, exact, fp32
    and treecode modes) equals the single-rank evaluation of the same steps bit for bit on every rank."""
    r = _torchrun(ngpus, ["scripts/dist_parity.py"])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("bitwise_equal_to_single_rank=True") == 5 * ngpus, r.stdout


def test_bench_line_parity_on_several_gpus(ngpus):
    """bench.py under torchrun at a reduced N: the JSON line's parity objects (self-convection vs oracle + sharded vs
    unsharded, flow-field slab vs oracle, sweep slice vs oracle) must all be ok."""
    r = _torchrun(ngpus, ["bench.py", "--gpus", str(ngpus), "--steps", "1", "--warmup", "1", "--nvortices", str(1 << 18)])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == ngpus
    assert line["parity"]["ok"] and line["parity"]["sharded_2steps_n131072_equals_unsharded_bitwise"] is True, line["parity"]
    assert line["flowfield"]["parity"]["ok"] and line["sweep"]["parity"]["ok"], (line["flowfield"]["parity"], line["sweep"]["parity"])
    assert line["config"]["transport"] in ("p2p", "nccl") and line["gpu_launches"] >= 1
    assert line["tree"]["parity"]["ok"] and line["tree"]["n_2p24"]["parity"]["ok"], line["tree"]
