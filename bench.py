#!/usr/bin/env python
"""bench.py -- users/sec of MultiVAE / MultiDAE training (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W [--config cfg2|cfg3|cfg4|cfg5]     # this repo's CUDA engine
    python bench.py --impl reference --steps K --warmup W                            # the reference's CPU path
    python bench.py --mode eval [--impl reference]                                   # evaluate(): users/sec

Workloads (BASELINE.json `configs`; synthetic Zipf / log-normal matrices from rectorch_b200/synth.py):
  cfg2 (default)  MultiVAE [50000-600-200], 200K users x 50K items per GPU, batch 500 per GPU, dropout 0.5,
                  beta 0.2 annealed over 20000 steps, Adam lr 1e-3
  cfg3            MultiDAE [50000-200], same matrix, lam 0.2, coupled weight decay 1e-3
  cfg4            cfg2's network on 1M users, batch 1000 per GPU
  cfg5            MultiVAE [200000-1024-512] on 2M users x 200K items, batch 256 per GPU (2048 on 8 GPUs)
One "step" = one train_batch (forward, loss, backward, Adam) on `batch` users per GPU.  Weak scaling: per-GPU batch
and per-GPU user shard are fixed, the global batch is batch * N; users are sharded row-wise and the gradients are
exchanged as described in DESIGN.md section 5 (reduce-scatter + sharded Adam + all-gather for W_d).

One JSON line on stdout (rank 0):
  value        users/s with the CSR matrix resident in HBM (CUDA-event timed, max over ranks)
  e2e          users/s through the public host-batch call (MultiVAE.train_batch_csr -> b200vae_train_step_host at
               N = 1): pinned host CSR batch H2D, the step, loss D2H + sync, every step inside the timed region
  roofline     dominant kernel of the step (by measured time) vs MEASURED_PEAKS.json
  roofline_k4  the fused decoder-GEMM + log-softmax kernel (north-star kernel): in-step and alone (CUDA-graph replay)
  cpu_baseline the unmodified reference (baseline/_ref) -- or, if that is absent, the oracle port -- on the host cores
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CONFIGS = {
    "cfg2": {"model": "vae", "dec_dims": [200, 600, 50000], "n_users": 200_000, "n_items": 50_000, "batch": 500,
             "dropout": 0.5, "beta": 0.2, "anneal_steps": 20000, "lam": 0.0, "lr": 1e-3, "arch": "MultiVAE [50000-600-200]"},
    "cfg3": {"model": "dae", "dec_dims": [200, 50000], "n_users": 200_000, "n_items": 50_000, "batch": 500,
             "dropout": 0.5, "beta": 0.0, "anneal_steps": 0, "lam": 0.2, "lr": 1e-3, "arch": "MultiDAE [50000-200]"},
    "cfg4": {"model": "vae", "dec_dims": [200, 600, 50000], "n_users": 1_000_000, "n_items": 50_000, "batch": 1000,
             "dropout": 0.5, "beta": 0.2, "anneal_steps": 20000, "lam": 0.0, "lr": 1e-3, "arch": "MultiVAE [50000-600-200]"},
    "cfg5": {"model": "vae", "dec_dims": [512, 1024, 200000], "n_users": 2_000_000, "n_items": 200_000, "batch": 256,
             "dropout": 0.5, "beta": 0.2, "anneal_steps": 20000, "lam": 0.0, "lr": 1e-3, "arch": "MultiVAE [200000-1024-512]"},
}
MAX_USERS_PER_RANK = 250_000      # bounds the host-side generation time (~10 s); stated in config.workload


def metric_name(cfg):
    return "users/sec (%s, batch %d/GPU, synthetic %dK users x %dK items)" % (
        cfg["arch"], cfg["batch"], cfg["n_users"] // 1000, cfg["n_items"] // 1000)


def eval_metric_name(cfg):
    return "users/sec of evaluate(recall@20, ndcg@100) (%s, batch %d, %dK items)" % (
        cfg["arch"], cfg["batch"], cfg["n_items"] // 1000)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


def load_ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the
    shipped build (profiles/r2_ncu_traffic.json, written by scripts/ncu_summary.py); empty when not captured."""
    p = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).  nvidia-smi needs 0.1-0.3 s
    to print its first sample, so it is started before a pre-roll of untimed steps (the GPU stays under load) and its
    time-stamped samples are filtered to the wall-clock window of the timed region afterwards."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    PERIOD_MS = 10

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", str(self.PERIOD_MS)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    @staticmethod
    def now():
        import datetime
        return datetime.datetime.now()

    def stop(self, t0=None, t1=None):
        """t0 / t1: ClockSampler.now() taken (after a device sync) at both ends of the timed region."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            self.proc.kill()
            out = ""
        return self.parse(out, t0, t1)

    @classmethod
    def parse(cls, out, t0=None, t1=None):
        """nvidia-smi csv lines -> the `clocks` object of the bench line; samples outside [t0, t1] (+- one period) are
        dropped when at least one falls inside."""
        import datetime
        rows = []
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f")
            except ValueError:
                ts = None
            try:
                rows.append((ts, float(f[1]), float(f[2]), [v.lower().startswith("active") for v in f[4:8]]))
            except ValueError:
                continue
        inside = rows
        window = "whole sampling period (pre-roll + timed region)"
        if t0 is not None and t1 is not None:
            slack = datetime.timedelta(milliseconds=cls.PERIOD_MS)
            sel = [r for r in rows if r[0] is not None and t0 - slack <= r[0] <= t1 + slack]
            if sel:
                inside, window = sel, "timed region"
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for r in inside for n, on in zip(names, r[3]) if on})
        sm = [r[1] for r in inside]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": inside[-1][2] if inside else None,
                "reasons": reasons, "samples": len(sm), "samples_total": len(rows), "window": window}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------
def pick_cpu_threads(n_items, hidden):
    """All the host threads the reference can actually use: start from the affinity mask / cgroup quota and
    keep the count that makes the reference's dominant op (the [B x H] x [H x n_items] sgemm) fastest --
    on shared hosts os.cpu_count() oversubscribes the container's CPU quota and is several times slower."""
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            avail = max(1, min(avail, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    cands = sorted({max(1, avail), max(1, avail // 2), max(1, avail // 4), min(avail, 32), min(avail, 16), min(avail, 8)})
    a = torch.randn(500, hidden)
    b = torch.randn(hidden, n_items)
    best, best_t = cands[-1], float("inf")
    for n in cands:
        torch.set_num_threads(n)
        torch.mm(a, b)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.mm(a, b)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def import_reference():
    """The UNMODIFIED reference from baseline/_ref (pip-installed copy of /root/reference; see baseline/README.md)
    with the two import shims it needs here (`bottleneck`, `munch` are not installable offline: oracle/_stubs).
    Returns the `rectorch` package or None."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "rectorch")):
        return None
    for p in (os.path.join(ROOT, "oracle", "_stubs"), ref):
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        import rectorch                   # noqa: F401
        import rectorch.models            # noqa: F401
        import rectorch.samplers          # noqa: F401
        import rectorch.evaluation        # noqa: F401
        return rectorch
    except Exception as e:                # noqa: BLE001
        sys.stderr.write("baseline/_ref import failed: %r\n" % (e,))
        return None


def cpu_reference_run(cfg, steps, warmup, budget_s, mode="train"):
    """Times the reference's own loop body on the host cores, on a bounded sample of the workload:
      kind "reference": rectorch.models.MultiVAE/MultiDAE.train_batch on the batches rectorch.samplers.DataSampler
                        yields (samplers.py:91-107 + models.py:817-835), i.e. the loop of train_epoch (models.py:409-410);
      kind "port":      oracle.train_step on the same dense batches (only when baseline/_ref is absent).
    mode "eval": rectorch.evaluation.evaluate(model, sampler, ["recall@20", "ndcg@100"]) (evaluation.py:67-110)."""
    from rectorch_b200 import synth
    B = cfg["batch"]
    H = cfg["dec_dims"][-2]
    cores = pick_cpu_threads(cfg["n_items"], H)
    total = max(steps + warmup, 2)
    n_rows = B * total if mode == "train" else B * max(steps, 4)
    csr = synth.make_matrix(n_rows, cfg["n_items"], seed=synth.DEFAULT_SEED)
    ref = import_reference()
    vae = cfg["model"] == "vae"
    torch.manual_seed(0)
    if ref is not None:
        kind = "reference"
        import rectorch.evaluation as reval
        import rectorch.models as rmodels
        import rectorch.nets as rnets
        import rectorch.samplers as rsamplers
        if vae:
            net = rnets.MultiVAE_net(list(cfg["dec_dims"]), None, cfg["dropout"])
            model = rmodels.MultiVAE(net, beta=cfg["beta"], anneal_steps=cfg["anneal_steps"], learning_rate=cfg["lr"])
        else:
            net = rnets.MultiDAE_net(list(cfg["dec_dims"]), None, cfg["dropout"])
            model = rmodels.MultiDAE(net, lam=cfg["lam"], learning_rate=cfg["lr"])
        if mode == "eval":
            tr, te = synth.split_heldout(csr, 0.2, seed=synth.DEFAULT_SEED + 1)
            sampler = rsamplers.DataSampler(tr.to_scipy(), te.to_scipy(), batch_size=B, shuffle=False)
            n_b = len(sampler)
            t0 = time.perf_counter()
            reval.evaluate(model, sampler, ["recall@20", "ndcg@100"])
            dt = time.perf_counter() - t0
            return {"value": n_b * B / dt, "ms_per_step": 1e3 * dt / n_b, "batch": B, "cores": cores, "kind": kind,
                    "sample": "rectorch.evaluation.evaluate on %d batches x %d users (predict + top-k + metrics), "
                              "torch %s CPU, %d threads" % (n_b, B, torch.__version__, cores)}
        sampler = rsamplers.DataSampler(csr.to_scipy(), batch_size=B, shuffle=False)
        model.network.train()
        times = []
        it = iter(sampler)
        t_budget = time.perf_counter()
        for i in range(total):
            t0 = time.perf_counter()
            data, gt = next(it)                    # samplers.py:99-105: CSR rows -> dense float tensor
            model.train_batch(data, gt)            # models.py:817-835 / 424-447
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if time.perf_counter() - t_budget > budget_s and len(times) >= 2:
                break
        what = "rectorch.samplers.DataSampler batch + rectorch.models.%s.train_batch (unmodified reference, baseline/_ref)" % (
            "MultiVAE" if vae else "MultiDAE")
    else:
        kind = "port"
        from oracle import multvae_oracle as O
        from rectorch_b200.nets import MultiDAE_net, MultiVAE_net
        sp = csr.to_scipy()
        net = (MultiVAE_net if vae else MultiDAE_net)(list(cfg["dec_dims"]), None, cfg["dropout"])
        onet = O.Net.from_state_dict({k: v.detach() for k, v in net.state_dict().items()}, vae, cfg["dropout"])
        ost = O.AdamState(onet, lr=cfg["lr"], weight_decay=0.0 if vae else 1e-3)
        if mode == "eval":
            tr, te = synth.split_heldout(csr, 0.2, seed=synth.DEFAULT_SEED + 1)
            t0 = time.perf_counter()
            O.evaluate(onet, tr.to_scipy(), te.to_scipy(), B, ["recall@20", "ndcg@100"])
            dt = time.perf_counter() - t0
            n_b = n_rows // B
            return {"value": n_rows / dt, "ms_per_step": 1e3 * dt / n_b, "batch": B, "cores": cores, "kind": kind,
                    "sample": "oracle.evaluate on %d batches x %d users, torch %s CPU, %d threads" % (n_b, B, torch.__version__, cores)}
        times = []
        t_budget = time.perf_counter()
        for i in range(total):
            t0 = time.perf_counter()
            x = torch.from_numpy(np.asarray(sp[i * B:(i + 1) * B].toarray(), dtype=np.float32))
            drop, eps = O.replay_rng_tape(100 + i, B, cfg["n_items"], cfg["dec_dims"][0], cfg["dropout"], vae)
            O.train_step(onet, ost, x, None, beta=O.beta_schedule(cfg["beta"], cfg["anneal_steps"], i) if vae else 0.0,
                         lam=cfg["lam"], drop_scale=drop, eps=eps)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
            if time.perf_counter() - t_budget > budget_s and len(times) >= 2:
                break
        what = "oracle port of the reference loop (dense batch + RNG + fwd + bwd + Adam)"
    ms = 1e3 * float(np.mean(times))
    return {"value": B / (ms / 1e3), "ms_per_step": ms, "batch": B, "cores": cores, "kind": kind,
            "sample": "%d steps x %d users of %s: %s, torch %s CPU ops, %d threads" % (
                len(times), B, cfg["arch"], what, torch.__version__, cores)}


def run_reference(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reference_run(cfg, args.steps, args.warmup, budget_s=170.0, mode=args.mode)
    line = {"impl": "reference", "metric": metric_name(cfg) if args.mode == "train" else eval_metric_name(cfg),
            "value": r["value"], "unit": "users/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %s, batch %d (bounded sample of the %dK x %dK matrix)" % (
                args.config, cfg["arch"], r["batch"], cfg["n_users"] // 1000, cfg["n_items"] // 1000),
                "global_batch": r["batch"], "timing": "host wall clock; inputs (the weights) exceed the CPU caches"},
            "cpu_baseline": {"value": r["value"], "unit": "users/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "users/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------
def build_model(cfg, dev):
    from rectorch_b200.models import MultiDAE, MultiVAE
    from rectorch_b200.nets import MultiDAE_net, MultiVAE_net
    torch.manual_seed(0)
    if cfg["model"] == "vae":
        net = MultiVAE_net(list(cfg["dec_dims"]), None, cfg["dropout"]).cuda(dev)
        return MultiVAE(net, beta=cfg["beta"], anneal_steps=cfg["anneal_steps"], learning_rate=cfg["lr"])
    net = MultiDAE_net(list(cfg["dec_dims"]), None, cfg["dropout"]).cuda(dev)
    return MultiDAE(net, lam=cfg["lam"], learning_rate=cfg["lr"])


def gather_global_matrix(local, world, dev):
    """Every rank generated its own shard of users (different seed); the replicated sampler wants the whole matrix on
    every rank: all-gather the CSR arrays over NCCL and stitch them in rank order."""
    import torch.distributed as dist
    from rectorch_b200 import synth
    n_loc, n_items = local.shape
    nnz = torch.tensor([local.nnz], dtype=torch.int64, device=dev)
    all_nnz = [torch.zeros_like(nnz) for _ in range(world)]
    dist.all_gather(all_nnz, nnz)
    all_nnz = [int(t.item()) for t in all_nnz]
    mx = max(all_nnz)
    idx = torch.zeros(mx, dtype=torch.int32, device=dev)
    idx[:local.nnz] = torch.from_numpy(local.indices).to(dev)
    idx_all = torch.empty(world * mx, dtype=torch.int32, device=dev)
    dist.all_gather_into_tensor(idx_all, idx)
    ip = torch.from_numpy(local.indptr).to(dev)
    ip_all = torch.empty(world * (n_loc + 1), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(ip_all, ip)
    idx_all = idx_all.cpu().numpy().reshape(world, mx)
    ip_all = ip_all.cpu().numpy().reshape(world, n_loc + 1)
    indptr = [np.zeros(1, dtype=np.int64)]
    base = 0
    for r in range(world):
        indptr.append(ip_all[r, 1:] + base)
        base += all_nnz[r]
    indices = np.concatenate([idx_all[r, :all_nnz[r]] for r in range(world)])
    return synth.CSR(np.concatenate(indptr), indices, np.ones(len(indices), np.float32), (n_loc * world, n_items))


def host_batches(csr, B, n):
    """n pinned host CSR batches of B consecutive users each."""
    out = []
    for b in range(n):
        sl = csr.rows(b * B, (b + 1) * B)
        out.append((torch.from_numpy(sl.indptr.copy()).pin_memory(), torch.from_numpy(sl.indices.copy()).pin_memory()))
    return out


def k4_alone(eng, model, cfg, dev, Bh, peaks):
    """The fused decoder GEMM + log-sum-exp kernel alone, launched back to back over rotating fp16 copies of W_d
    (so that no launch finds its weights in L2) from ONE CUDA graph: per-launch time without host launch cost."""
    from rectorch_b200 import _lib
    from rectorch_b200._lib import check
    I, H = cfg["n_items"], cfg["dec_dims"][-2]
    Wd = model.network.dec_layers[-1].weight.detach().half().contiguous()
    bd = model.network.dec_layers[-1].bias.detach()
    ncopy = max(3, int(400e6 // (2 * I * H)) + 1)
    copies = [Wd.clone() for _ in range(ncopy)]
    hh = torch.tanh(torch.randn(Bh, H, device=dev)).half()
    reps = 4 * ncopy
    side = torch.cuda.Stream(device=dev)

    def call(k, sp):
        check(_lib.lib().b200vae_dec_fwd_lse(eng._ctx, ctypes.c_void_p(hh.data_ptr()),
                                             ctypes.c_void_p(copies[k % ncopy].data_ptr()),
                                             ctypes.c_void_p(bd.data_ptr()), Bh, I, H, None, sp))
    with torch.cuda.stream(side):
        for k in range(ncopy):
            call(k, ctypes.c_void_p(side.cuda_stream))
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        sp = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        for k in range(reps):
            call(k, sp)
    g.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / (3 * reps)
    byt = 2.0 * I * H + 4.0 * I + 2.0 * Bh * H + 16.0 * Bh * 74
    fl = 2.0 * Bh * I * H
    del copies
    return {"batch": Bh, "ms": ms, "gbs": byt / ms / 1e6, "frac": byt / ms / 1e6 / peaks["hbm_gbs"],
            "tflops": fl / ms / 1e9, "tensor_frac": fl / ms / 1e9 / peaks["bf16_tflops"],
            "algorithmic_bytes": byt, "launches": 3 * reps,
            "how": "kernel alone, %d launches captured in one CUDA graph (no host launch cost) over %d rotating fp16 "
                   "copies of W_d (%d MB >> L2)" % (reps, ncopy, ncopy * 2 * I * H // 1000000)}


def run_b200(args, cfg):
    import torch.distributed as dist
    from rectorch_b200 import synth
    from rectorch_b200.samplers import DataSampler

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world == 1 and args.gpus > 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    B, K, W = cfg["batch"], args.steps, args.warmup
    I, H = cfg["n_items"], cfg["dec_dims"][-2]
    n_users = min(cfg["n_users"] // (world if args.config in ("cfg4", "cfg5") else 1), MAX_USERS_PER_RANK)
    model = build_model(cfg, dev)
    eng = model._engine
    t0 = time.perf_counter()
    shard = synth.make_matrix(n_users, I, seed=synth.DEFAULT_SEED + rank)
    if world == 1:
        csr = shard
        sampler = DataSampler(csr, None, batch_size=B, shuffle=False, device=dev)
    else:
        # one global matrix of n_users x world rows, users sharded row-wise: rank r trains on rows
        # [r * n_users, (r + 1) * n_users), the global batch is B * world
        csr = gather_global_matrix(shard, world, dev)
        sampler = DataSampler(csr, None, batch_size=B * world, shuffle=False, device=dev, rank=rank, world_size=world,
                              replicate=True)
    gen_s = time.perf_counter() - t0
    batches = list(sampler.iter_rows(dev))
    model.network.train()
    slots = model._loss_hist

    def step(i):
        rb = batches[i % len(batches)]
        beta, lam = model._step_coeffs()
        model._step(rb, None, beta, lam, slots[4 * (i % 1024):4 * (i % 1024) + 4])
        model._after_step()

    def barrier():
        if world > 1:
            torch.cuda.current_stream(dev).wait_stream(model._comm_stream)
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(W):
        step(i)
    barrier()
    # pre-roll: ~0.3 s of untimed steps with nvidia-smi already running (it needs 50-300 ms to print its first sample),
    # so that its samples cover the timed region.  The step time is estimated on 5 steady-state steps AFTER the
    # warm-up (the first warm-up step pays for context creation), and the count is the same on every rank: it is
    # derived from the slowest rank's estimate.
    t_w = time.perf_counter()
    for i in range(5):
        step(W + K + i)
    barrier()
    est = torch.tensor([(time.perf_counter() - t_w) / 5.0], device=dev)
    if world > 1:
        dist.all_reduce(est, op=dist.ReduceOp.MAX)
    n_pre = int(min(2000, max(50, 0.3 / max(float(est.item()), 1e-5))))
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for i in range(n_pre):
        step(W + K + i)
    barrier()
    eng.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_region0 = ClockSampler.now()
    ev0.record()
    for i in range(W, W + K):
        step(i)
    if world > 1:
        torch.cuda.current_stream(dev).wait_stream(model._comm_stream)
    ev1.record()
    barrier()
    t_region1 = ClockSampler.now()
    ms_total = ev0.elapsed_time(ev1)
    clk = clocks.stop(t_region0, t_region1) if rank == 0 else None
    launches = eng.launch_count()
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    eng.check_overflow()
    last_loss = float(slots[4 * ((W + K - 1) % 1024)].item())
    ms_step = ms_total / K
    value = world * B / (ms_step / 1e3)

    # ---- host cost of issuing one step (python + ctypes + launches), GPU idle at the start, no sync inside ----
    barrier()
    t0 = time.perf_counter()
    for i in range(20):
        step(W + K + i)
    host_us = (time.perf_counter() - t0) / 20 * 1e6
    barrier()

    # ---- e2e: pinned host CSR batches through the public host-batch call, H2D + step + loss D2H every step ----
    nb = 32
    e2e_steps = min(K, 200)
    if world == 1:
        host = host_batches(csr, B, nb)
    else:
        # every rank passes the GLOBAL batch: rank-major concatenation of the ranks' next B users
        host = []
        for b in range(nb):
            parts = [csr.rows(r * n_users + b * B, r * n_users + (b + 1) * B) for r in range(world)]
            offs = np.cumsum([0] + [p.nnz for p in parts])
            ip = np.concatenate([np.zeros(1, np.int64)] + [p.indptr[1:] + offs[j] for j, p in enumerate(parts)])
            ix = np.concatenate([p.indices for p in parts])
            host.append((torch.from_numpy(ip).pin_memory(), torch.from_numpy(ix).pin_memory()))
    h2d = float(np.mean([ip.numel() * 8 + ix.numel() * 4 for ip, ix in host]))
    for i in range(3):
        model.train_batch_csr(*host[i % nb])
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        model.train_batch_csr(*host[i % nb])
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e = {"value": world * B * e2e_steps / dt, "unit": "users/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16,
           "steps": e2e_steps,
           "api": "MultiVAE.train_batch_csr(pinned host CSR batch) -> loss float; " + (
               "one call of b200vae_train_step_host (H2D, step, loss D2H, stream sync)" if world == 1 else
               "every rank copies the global batch's CSR (H2D), runs the data-parallel step, reads the loss back")}
    if world == 1 and cfg["n_items"] <= 50000:
        # the reference-shaped call: train_batch(dense host FloatTensor [B x I]) -- PCIe bound (100 MB/step)
        xd = torch.from_numpy(csr.rows(0, B).toarray()).pin_memory()
        model.train_batch(xd)
        barrier()
        t0 = time.perf_counter()
        for i in range(10):
            model.train_batch(xd)
        barrier()
        dtd = time.perf_counter() - t0
        e2e["dense_api"] = {"value": B * 10 / dtd, "unit": "users/s", "h2d_bytes_per_step": int(xd.numel() * 4),
                            "api": "train_batch(pinned dense FloatTensor) as the reference's loop does"}

    # ---- data-parallel parity: the same global batch through the N-rank step and through ONE process ----
    dp_parity = None
    if world > 1:
        model._bind_sampler(sampler)
        dp_parity = dp_parity_check(model, sampler, n_users, B, world, rank, dev)

    # ---- phases of the data-parallel step (CUDA events on both streams; not part of the timed region above) ----
    dp_phases = None
    if world > 1 and args.dp_timing:
        from rectorch_b200.models import _PhaseTimer
        model._dp_timer = _PhaseTimer()
        for i in range(10):
            step(W + K + 100 + i)
        barrier()
        dp_phases = [{"phase": k, "done_at_us": v} for k, v in model._dp_timer.report()]
        model._dp_timer = None

    # ---- per-kernel timing pass (CUDA events inside the library, on the launching stream; serial schedule) ----
    eng.set_timing(True)
    names = ["dec_fwd_lse(K4)", "adam(K8)", "dec_bwd_prob(K5)", "dWd_gemm", "dh_gemm"]
    acc = np.zeros(5)
    reps = 20
    for i in range(reps):
        step(W + K + i)
        barrier()
        acc += np.array([eng.kernel_ms(j) for j in range(5)])
    eng.set_timing(False)
    kms = acc / reps
    peaks = load_peaks()
    traffic = load_ncu_traffic().get(args.config, {})
    P = sum(int(np.prod(s)) + s[0] for s in eng.shapes)          # parameters (28 B each through Adam: SURVEY 8d)
    adam_bytes = 28.0 * P
    k4_bytes = 2.0 * I * H + 4.0 * I + 2.0 * B * H + 16.0 * B * 74     # fp16 W_d image + bias + h + (max,sum) partials
    k4_flops = 2.0 * B * I * H
    k4_gbs = k4_bytes / (kms[0] * 1e-3) / 1e9 if kms[0] > 0 else None
    k4_tf = k4_flops / (kms[0] * 1e-3) / 1e12 if kms[0] > 0 else None
    roof_dom = None
    if world == 1:
        dom = int(np.argmax(kms))
        if dom == 1:
            ach = adam_bytes / (kms[1] * 1e-3) / 1e9
            roof_dom = {"kernel": names[1], "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                        "frac": ach / peaks["hbm_gbs"], "traffic": traffic.get("adam"), "algorithmic_bytes": adam_bytes,
                        "ms": float(kms[1]), "peak_src": peaks["src"] + " (copy bandwidth, MEASURED_PEAKS.json)",
                        "note": "28 B/param (w,g,m,v read; w,m,v written) x %d params; the kernel also writes the fp16 image of "
                                "W_d (%.0f MB, not counted)" % (P, 2.0 * I * H / 1e6)}
        else:
            flops = {0: k4_flops, 2: k4_flops, 3: 2.0 * B * I * (H + 8), 4: 2.0 * B * I * H}[dom]
            ach = flops / (kms[dom] * 1e-3) / 1e12
            pk = peaks["bf16_tflops_sustained"]
            roof_dom = {"kernel": names[dom], "bound": "tensor", "achieved": ach, "peak": pk, "unit": "TFLOP/s",
                        "frac": ach / pk, "traffic": traffic.get(names[dom]), "ms": float(kms[dom]),
                        "peak_src": peaks["src"] + " bf16 sustained (fp16 operands run at the bf16 rate)"}
    else:
        roof_dom = {"kernel": names[0], "bound": "tensor", "achieved": k4_tf, "peak": peaks["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": (k4_tf or 0) / peaks["bf16_tflops_sustained"], "traffic": traffic.get("k4"),
                    "ms": float(kms[0]), "note": "N > 1: Adam is sharded / split over streams, the N = 1 line carries its roofline"}
    k4_hbm = k4_b = None
    if world == 1 and eng.use_tc:
        k4_hbm = k4_alone(eng, model, cfg, dev, min(250, B), peaks)
        k4_b = k4_alone(eng, model, cfg, dev, B, peaks)
    roof_k4 = {"kernel": names[0], "ms_in_step": float(kms[0]), "hbm_gbs": k4_gbs, "hbm_frac": (k4_gbs or 0) / peaks["hbm_gbs"],
               "tflops_f16": k4_tf, "tensor_frac_of_bf16_burst": (k4_tf or 0) / peaks["bf16_tflops"],
               "algorithmic_bytes": k4_bytes, "flops": k4_flops, "traffic": traffic.get("k4"),
               "hbm_regime": k4_hbm, "at_batch": k4_b,
               "note": "ms_in_step = CUDA events around the tcgen05 GEMM + log-sum-exp kernel inside a training step (includes "
                       "~5 us of event overhead); the kernel reads the fp16 image of W_d (2 B/weight: half of SURVEY 8d's "
                       "fp32 figure) once; hbm_regime / at_batch = the kernel alone, graph-replayed, where it is HBM-bound "
                       "(B <= 250) and at the benchmark batch"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_run(cfg, 6, 2, budget_s=25.0)
        cpu = {"value": r["value"], "unit": "users/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}

    if rank == 0:
        par = "dp1" if world == 1 else (
            "dp%d row-sharded; W_d: reduce-scatter + 1/N Adam + all-gather of the fp16 image; encoder-0 gradient from "
            "all-gathered factors; hidden layers + b_d: one small all-reduce" % world if model._zero else
            "dp%d row-sharded; all-reduce of dW_d + encoder-0 gradient from all-gathered factors" % world)
        line = {"metric": metric_name(cfg), "value": value, "unit": "users/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32 (item-sized GEMMs: fp16 tensor-core operands with 10-bit mantissa, fp32 accumulate; fp32 master weights)",
                "data": "synthetic",
                "config": {"workload": "%s: %s, %d users x %d items per GPU (%.1f s to generate), batch %d per GPU" % (
                    args.config, cfg["arch"], n_users, I, gen_s, B), "global_batch": B * world, "parallelism": par,
                    "schedule": "decoder-output Adam on a second stream beside the encoder backward" if world == 1 else
                                "W_d exchange + sharded Adam on a side stream, overlapped with the encoder backward and the next step's encoder",
                    "l2": "no flush: per-step working set (4 x %d MB arenas) exceeds the 126 MB L2" % (4 * P // 1000000),
                    "pre_roll_steps": int(n_pre)},
                "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "host_issue_us_per_step": host_us,
                "roofline": roof_dom, "roofline_k4": roof_k4,
                "kernel_ms": {n: float(v) for n, v in zip(names, kms)},
                "cpu_baseline": cpu, "last_loss": last_loss, "dp_parity": dp_parity, "dp_phases": dp_phases}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def dp_parity_check(model, sampler, n_users, B, world, rank, dev):
    """One more data-parallel step on a fresh global batch, replayed by rank 0 alone (one process, batch B * N,
    same weights, same Philox seed): loss and the updated weights must agree up to summation order.
    Lets a multi-GPU bench run verify the sharded path that a 1-GPU test box cannot."""
    import torch.distributed as dist
    from rectorch_b200.samplers import RowBatch
    eng = model._engine
    model.sync_weights()
    torch.cuda.synchronize(dev)
    w0, m0, v0 = eng.w.clone(), eng.m.clone(), eng.v.clone()
    steps0 = eng.adam_steps
    b0 = 40       # a batch index; any would do
    rows_g = torch.from_numpy(np.concatenate([r * n_users + np.arange(b0 * B, (b0 + 1) * B)
                                              for r in range(world)]).astype(np.int32)).to(dev)
    rb = RowBatch(sampler, rows_g[rank * B:(rank + 1) * B].contiguous(), False, rows_g)
    beta, lam = model._step_coeffs()
    seed = (model._dp_seed + 0x9E3779B97F4A7C15 * (eng.adam_steps + 1)) % (1 << 63)
    slot = model._loss_hist[:4]
    model._step(rb, None, beta, lam, slot)
    model.sync_weights()
    torch.cuda.current_stream(dev).wait_stream(model._comm_stream)
    torch.cuda.synchronize(dev)
    loss_dp = model._loss_from(slot, beta, lam)
    w_dp = eng.w.clone()
    out = None
    if rank == 0:
        # single-process replay on the same engine: restore the state, run the global batch as ONE local batch
        sharded = eng._w1_shard
        if sharded:
            eng.set_w1_sharding(1, 0)           # the replay is an ordinary single-process step
        eng.w.copy_(w0)
        eng.m.copy_(m0)
        eng.v.copy_(v0)
        eng.adam_steps = steps0
        lr, betas, eps, wd = model._hyper()
        eng.loss_buf = torch.zeros(4, dtype=torch.float32, device=dev)
        eng.train_step(rows=rows_g, beta=beta, lam=lam, dropout_p=float(model.network.dropout.p), seed=seed, lr=lr,
                       betas=betas, eps=eps, weight_decay=wd)
        torch.cuda.synchronize(dev)
        loss_1 = float(eng.loss_buf[0].item())
        d = (eng.w - w_dp).abs()
        if sharded:
            eng.w.copy_(w_dp)
            eng.set_w1_sharding(*sharded)
        out = {"loss_dp": loss_dp, "loss_single_process": loss_1, "loss_rel": abs(loss_dp - loss_1) / abs(loss_1),
               "w_max_abs_diff": float(d.max().item()), "w_frac_diff_gt_1e-5": float((d > 1e-5).float().mean().item()),
               "how": "global batch of %d users: %d-rank step vs the same batch in one process on rank 0 "
                      "(same weights, same Philox seed)" % (B * world, world)}
    # everybody adopts rank 0's copy of the data-parallel result so that the replicas stay identical
    eng.w.copy_(w_dp)
    eng.adam_steps = steps0 + 1
    dist.barrier()
    model._broadcast_state()
    return out


def run_eval(args, cfg):
    """evaluate(model, sampler, ["recall@20", "ndcg@100"]) throughput: predict (eval forward + seen-item mask) +
    device top-k + metric reduction per batch (rectorch/evaluation.py:67-110)."""
    from rectorch_b200 import synth
    from rectorch_b200.evaluation import evaluate
    from rectorch_b200.samplers import DataSampler
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    B, I = cfg["batch"], cfg["n_items"]
    n_b = max(args.steps, 4)
    csr = synth.make_matrix(B * n_b, I, seed=synth.DEFAULT_SEED)
    tr, te = synth.split_heldout(csr, 0.2, seed=synth.DEFAULT_SEED + 1)
    model = build_model(cfg, dev)
    sampler = DataSampler(tr, te, batch_size=B, shuffle=False, device=dev)
    mets = ["recall@20", "ndcg@100"]
    for _ in range(max(1, args.warmup // 3)):
        evaluate(model, sampler, mets)
    torch.cuda.synchronize(dev)
    clocks = ClockSampler(dev.index or 0)
    clocks.start()
    t_w = time.perf_counter()
    evaluate(model, sampler, mets)
    torch.cuda.synchronize(dev)
    per_eval = max(time.perf_counter() - t_w, 1e-4)
    for _ in range(int(min(200, 0.3 / per_eval))):          # pre-roll under load while nvidia-smi starts up
        evaluate(model, sampler, mets)
    reps = int(max(3, min(100, 0.1 / per_eval)))            # a timed region of >= ~0.1 s
    torch.cuda.synchronize(dev)
    model._engine.launch_count(reset=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_region0 = ClockSampler.now()
    ev0.record()
    for _ in range(reps):
        res = evaluate(model, sampler, mets)
    ev1.record()
    torch.cuda.synchronize(dev)
    t_region1 = ClockSampler.now()
    ms = ev0.elapsed_time(ev1) / (reps * n_b)
    launches = model._engine.launch_count()
    clk = clocks.stop(t_region0, t_region1)
    t0 = time.perf_counter()
    for _ in range(reps):
        evaluate(model, sampler, mets)
    torch.cuda.synchronize(dev)
    wall = (time.perf_counter() - t0) / (reps * n_b)
    peaks = load_peaks()
    byt = 4.0 * B * I           # the [B x n_items] score matrix read once by the top-k (if materialised)
    r = cpu_reference_run(cfg, 4, 0, 30.0, mode="eval") if not args.no_cpu_baseline else None
    line = {"metric": eval_metric_name(cfg), "value": B / (ms / 1e3), "unit": "users/s", "n_gpus": 1, "steps": reps * n_b,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (fp16 tensor-core operands, fp32 accumulate)", "data": "synthetic",
            "config": {"workload": "%s eval: %s, %d held-out users, batch %d, recall@20 + ndcg@100" % (
                args.config, cfg["arch"], B * n_b, B)},
            "clocks": clk,
            "e2e": {"value": B / wall, "unit": "users/s", "h2d_bytes_per_step": B * 4, "d2h_bytes_per_step": 2 * B * 4,
                    "api": "evaluation.evaluate(model, DataSampler(tr, te)): wall clock incl. the host loop and the "
                           "device -> host read of the per-user metric vectors"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "predict + top-k (per batch)", "bound": "hbm", "achieved": byt / ms / 1e6,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": byt / ms / 1e6 / peaks["hbm_gbs"], "traffic": None,
                         "note": "4*B*I bytes (one pass over the score matrix) / time of the whole batch"},
            "cpu_baseline": None if r is None else {"value": r["value"], "unit": "users/s", "cores": r["cores"],
                                                    "kind": r["kind"], "sample": r["sample"]},
            "metrics": {m: float(np.nanmean(v)) for m, v in res.items()}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--mode", default="train", choices=["train", "eval"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dp-timing", action="store_true", help="N > 1: also report when each phase of the step completes")
    ap.add_argument("--dp", default="zero", choices=["zero", "factors"],
                    help="N > 1: gradient exchange (zero: sharded Adam for W_d; factors: all-reduce of dW_d)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.dp == "factors":
        os.environ["B200VAE_DP_ZERO"] = "0"
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg)
    elif args.mode == "eval":
        run_eval(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == "__main__":
    main()
