"""CPU: host-side logic, the C-ABI surface (load + exported symbols, no compute), and the
world_size-2 gloo check of the data-parallel plumbing."""
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from rectorch_b200 import _lib, synth
from rectorch_b200.metrics import KINDS, parse_metric
from rectorch_b200.samplers import DataSampler, shard_plan
from tests._util import load_golden, state_dict_from

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b200vae.h")).read()
    declared = set(re.findall(r"\b(b200vae_[a-z0-9_]+)\s*\(", hdr))
    declared.discard("b200vae_ctx")
    handle = _lib.lib()
    for name in sorted(declared):
        assert hasattr(handle, name), "libb200vae.so does not export %s" % name
    # and the ctypes table covers the header
    missing = declared - set(_lib.EXPORTS)
    assert not missing, "ctypes signatures missing for %s" % sorted(missing)
    assert handle.b200vae_version() >= 100


def test_no_cpu_fallback():
    """The product path must fail loudly without a GPU."""
    from rectorch_b200.models import MultiVAE
    from rectorch_b200.nets import MultiVAE_net
    net = MultiVAE_net([2, 4, 10])
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        MultiVAE(net)
    with pytest.raises(RuntimeError):
        net(torch.ones(2, 10))


@pytest.mark.parametrize("name", ["cfg1_dae", "small_vae", "vae_1layer", "small_dae"])
def test_init_matches_reference_bitwise(name):
    """Same construction order as rectorch/nets.py:208-216, 390 -> identical init for a seed."""
    from rectorch_b200.nets import MultiDAE_net, MultiVAE_net
    g = load_golden(name)
    torch.manual_seed(g["seed_net"])
    net = (MultiVAE_net if g["vae"] else MultiDAE_net)(list(g["dec_dims"]), None, g["p"])
    ref = state_dict_from(g, "init")
    sd = net.state_dict()
    assert list(sd.keys()) == list(ref.keys())
    for k in sd:
        assert sd[k].shape == ref[k].shape and torch.equal(sd[k], ref[k]), k


def test_net_attributes():
    """rectorch/tests/test_nets.py:40-46, 61-67 (the parts that need no forward pass)."""
    from rectorch_b200.nets import AE_net, MultiDAE_net, MultiVAE_net
    for cls in (MultiDAE_net, MultiVAE_net):
        net = cls([1, 2], [2, 1], .1)
        for a in ("enc_dims", "dec_dims", "dropout", "dec_layers", "enc_layers"):
            assert hasattr(net, a)
        assert isinstance(net.dropout, torch.nn.Dropout) and net.dropout.p == .1
    net = AE_net([1, 2], [2, 1])
    with pytest.raises(NotImplementedError):
        net.encode(torch.ones(1, 2))
    with pytest.raises(NotImplementedError):
        net.init_weights()
    assert AE_net([1, 2]).enc_dims == [2, 1]
    assert MultiVAE_net([3, 5, 7]).enc_layers[-1].out_features == 6     # 2 * latent


def test_shard_plan():
    for n, b, w in [(1000, 32, 2), (1001, 100, 4), (7, 4, 2), (10 ** 6, 1000, 8)]:
        plans = [shard_plan(n, b, r, w) for r in range(w)]
        assert plans[0][0] == 0 and plans[-1][1] == n
        for a, c in zip(plans, plans[1:]):
            assert a[1] == c[0]                      # contiguous cover
        assert len({p[2] for p in plans}) == 1 and len({p[3] for p in plans}) == 1
        assert all(p[1] - p[0] >= p[4] for p in plans)
        lb, nb, used = plans[0][2:]
        assert lb * w == b and nb == int(np.ceil(used / lb))
    with pytest.raises(ValueError):
        shard_plan(10, 5, 0, 2)
    with pytest.raises(ValueError):
        shard_plan(10, 4, 2, 2)


def test_sampler_host_side():
    m = synth.make_matrix(103, 50, seed=3, density=0.2)
    s = DataSampler(m, batch_size=10, shuffle=False)
    assert len(s) == 11 and s.n_users == 103 and s.n_items == 50
    s2 = DataSampler(m, batch_size=10, shuffle=True, rank=1, world_size=2)
    assert len(s2) == int(np.ceil((103 // 2) / 5)) and s2.row_offset == 0 and s2._lo == 51 and s2.replicate
    np.random.seed(5)
    a = s._permutation()
    assert np.array_equal(a, np.arange(103))
    s.shuffle = True
    np.random.seed(5)
    p1 = s._permutation()
    np.random.seed(5)
    ref = list(range(103))
    np.random.shuffle(ref)                      # what the reference does (samplers.py:93-95)
    assert np.array_equal(p1, np.array(ref))


def test_metric_parser():
    assert parse_metric("ndcg@10") == (KINDS["ndcg"], 10)
    assert parse_metric("Recall@20") == (KINDS["recall"], 20)
    assert parse_metric("hit_at_k") == (KINDS["hit"], 100)
    assert parse_metric("precision@10") is None and parse_metric("precision_at_k") is None


def test_synth_generator():
    m = synth.make_matrix(2000, 5000, seed=1)
    lens = np.diff(m.indptr)
    assert m.shape == (2000, 5000) and lens.min() >= 1 and m.nnz == lens.sum()
    for r in (0, 17, 1999):
        cols = m.indices[m.indptr[r]:m.indptr[r + 1]]
        assert np.all(np.diff(cols) > 0)         # sorted, no duplicates
    m2 = synth.make_matrix(2000, 5000, seed=1)
    assert np.array_equal(m.indices, m2.indices)
    tr, te = synth.split_heldout(m)
    assert tr.nnz + te.nnz == m.nnz
    assert np.array_equal(np.diff(tr.indptr) + np.diff(te.indptr), lens)
    d = m.rows(5, 9).toarray()
    assert d.shape == (4, 5000) and d.sum() == lens[5:9].sum()
    # popularity is Zipf-like: the head item is far more frequent than the median item
    cnt = np.bincount(m.indices, minlength=5000)
    assert cnt[0] > 20 * max(1, np.median(cnt))


_GLOO_WORKER = r"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %(root)r)
from rectorch_b200.samplers import shard_plan
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
# the step's exchange: local gradients already scaled by 1/B_global, ONE all_reduce(sum) on the flat arena,
# loss components summed the same way (rectorch_b200/models.py::AETrainer._step)
n_users, bg = 1001, 64
lo, hi, lb, nb, used = shard_plan(n_users, bg, rank, world)
rng = np.random.default_rng(0)
per_user = rng.standard_normal((n_users, 37)).astype(np.float32)      # stand-in per-user gradient rows
rows = np.arange(lo, hi)[:lb]
g = torch.from_numpy(per_user[rows].sum(0) / bg)
loss = torch.tensor([float(per_user[rows, 0].sum() / bg), 0., 0., 3.0])
dist.all_reduce(g); dist.all_reduce(loss)
all_rows = np.concatenate([np.arange(*shard_plan(n_users, bg, r, world)[:2])[:lb] for r in range(world)])
ref = per_user[all_rows].sum(0) / bg
assert np.allclose(g.numpy(), ref, atol=1e-5), "allreduced gradient != single-process gradient"
assert abs(loss[0].item() - per_user[all_rows, 0].sum() / bg) < 1e-4
assert abs(loss[3].item() / world - 3.0) < 1e-6          # replicated term is divided by world
# replicated sampler: with shuffling, rank 0's per-shard permutations reach every rank (broadcast), so all ranks
# agree on the rows of the whole global batch (RowBatch.all_rows) although their numpy generators differ
from rectorch_b200 import synth
from rectorch_b200.samplers import DataSampler
np.random.seed(100 + rank)
m = synth.make_matrix(n_users, 50, seed=2, density=0.2)
s = DataSampler(m, None, batch_size=bg, shuffle=True, rank=rank, world_size=world)
assert s.replicate and s.row_offset == 0 and len(s) == nb
plan = s.global_plan()
assert plan.shape == (world, used) and plan.dtype == np.int32
for r in range(world):
    l, h = shard_plan(n_users, bg, r, world)[:2]
    assert plan[r].min() >= l and plan[r].max() < h and len(set(plan[r].tolist())) == used
mine = torch.from_numpy(plan.astype(np.int64).copy())
lst = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(lst, mine)
assert all(torch.equal(lst[0], t) for t in lst), "ranks disagree on the global plan"
s0 = DataSampler(m, None, batch_size=bg, shuffle=False, rank=rank, world_size=world)
p0 = s0.global_plan()
assert np.array_equal(p0[1], np.arange(*shard_plan(n_users, bg, 1, world)[:2])[:used])
s1 = DataSampler(m, None, batch_size=bg, shuffle=False, rank=rank, world_size=world, replicate=False)
assert not s1.replicate and s1.row_offset == lo
# sharded optimiser (AETrainer._step_dp_zero): sum over ranks of the local W_d gradients, each rank updates only its
# shard and the shards are gathered back == the replicated update on the all-reduced gradient
from rectorch_b200.models import zero_shard
n_wd = 64 * world * 5
a, b = zero_shard(n_wd, world, rank)
assert (b - a) * world == n_wd and a == rank * (b - a) and zero_shard(n_wd + 8, world, rank) is None
gl = torch.from_numpy(np.random.default_rng(7 + rank).standard_normal(n_wd).astype(np.float32))
full = gl.clone(); dist.all_reduce(full)
parts = [torch.empty(b - a) for _ in range(world)]
for r in range(world):      # reduce_scatter emulated with reduce (gloo has no reduce_scatter)
    t = gl[r * (b - a):(r + 1) * (b - a)].clone(); dist.reduce(t, dst=r)
    if r == rank: mine_g = t
w_new = -0.001 * torch.sign(mine_g)                   # stand-in for Adam on the shard
lst = [torch.empty(b - a) for _ in range(world)]
dist.all_gather(lst, w_new)
assert torch.equal(torch.cat(lst), -0.001 * torch.sign(full)), "sharded update != replicated update"
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_gloo_world2_exchange(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER % {"root": ROOT})
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
