"""Multi-GPU parity of the fused peer-memory gradient exchange (csrc/dp.cu, parallel.PeerAdam): needs >= 2 GPUs on the
box (skipped otherwise; run with `gpurun --gpus 2 -- python -m pytest tests -m gpu`).  Replaces the replica gradient
SUM + Adam of /root/reference/sagan/main.py:190,205.  The check itself is tools/dp_check.py, run under torchrun."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_dp_check(world, port=29611):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "dp_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert lines, res.stdout[-2000:] + res.stderr[-2000:]
    return json.loads(lines[-1])


@pytest.mark.parametrize("world", [2])
def test_p2p_exchange_is_bit_identical_to_nccl_and_replicas_agree(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, box has {torch.cuda.device_count()}")
    out = run_dp_check(world)
    print(json.dumps(out))
    # one update through the peer-memory kernel == NCCL all-reduce + Adam, bit for bit at W = 2 (one addition per element)
    assert out["one_update_delta_max_abs"] == 0.0
    # after 3 full training steps every replica holds identical weights and reports identical (global) losses
    for mode in ("nccl", "p2p"):
        assert out[f"{mode}_G_replica_max_abs_diff"] == 0.0 and out[f"{mode}_D_replica_max_abs_diff"] == 0.0
        assert out[f"{mode}_loss_replica_max_abs_diff"] == 0.0
    assert out["p2p_vs_nccl_G_rel_l2_after_3_steps"] < 2e-3 and out["p2p_vs_nccl_D_rel_l2_after_3_steps"] < 2e-3
    assert out["replicas_with_own_noise"] == world
