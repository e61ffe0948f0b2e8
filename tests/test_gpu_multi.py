"""GPU (>= 2 devices): row-sharded data parallelism == single process on the same global batches.

Rank r owns users [r*U/N, (r+1)*U/N); every step each rank runs forward/backward on its local
rows with the loss scaled by 1/B_global and applies the identical fused Adam to the summed gradient.
"allreduce": sharded CSR, the whole gradient arena is all-reduced.  "factors" (default with a replicated
sampler): only the decoder-output half is all-reduced, the encoder-0 gradient is rebuilt on every rank from
the all-gathered delta rows (AETrainer._step_dp_factors).  "zero" (default): as "factors", but dW_d is reduce-scattered, every rank
runs Adam on its 1/N shard of W_d and the ranks all-gather the fp16 image (AETrainer._step_dp_zero); the fp32 weight and
its Adam moments are gathered back by sync_weights() before they are compared; encoder layer 0 is sharded the same way by
item rows j % N == rank ("zero_wd" leaves it replicated).  Dropout / eps come from Philox keyed by the GLOBAL user row, so the result
must equal the 1-process run up to fp32 summation order.
"""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("mode", ["zero", "zero_wd", "factors", "allreduce"])
@pytest.mark.parametrize("world", [2])
def test_sharded_training_matches_single_process(tmp_path, world, mode):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    worker = os.path.join(ROOT, "tests", "_mp_worker.py")
    multi, single = str(tmp_path / "multi.npz"), str(tmp_path / "single.npz")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), worker, multi, str(world), mode],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    env = dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, worker, single, str(world), mode], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    a, b = np.load(multi), np.load(single)
    rel = np.abs(a["losses"] - b["losses"]) / np.abs(b["losses"])
    # "zero": the encoder-0 rows the forward pass gathers are fp16 images (10-bit mantissa) of the sharded fp32
    # weights, so the run differs from the single-process one like any tensor-core operand rounding does
    tol = 1e-4 if mode == "zero" else 2e-6
    print("mode %s: max loss rel diff %.2e" % (mode, rel.max()))
    assert rel.max() < tol, "losses: multi %s single %s" % (a["losses"], b["losses"])
    for k in b.files:
        if k == "losses" or k.startswith("adam_"):
            continue
        d = np.abs(a[k] - b[k])
        # Adam normalises updates: a rounding-level gradient difference can move isolated weights by O(lr)
        lim = 1e-4 if mode == "zero" else 1e-5
        assert np.mean(d > lim) < 0.01, "%s: %.4f of the weights differ by > %.0e (max %.2e)" % (k, np.mean(d > lim), lim, d.max())
    # the sharded Adam moments of W_d, gathered back: same statistics as the replicated run
    for k in ("adam_m_wd", "adam_v_wd", "adam_m_w1", "adam_v_w1"):
        d = np.abs(a[k] - b[k])
        assert d.max() <= 1e-6 + (2e-2 if mode == "zero" else 1e-3) * np.abs(b[k]).max(), "%s: max diff %.3e" % (k, d.max())
