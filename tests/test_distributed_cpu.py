"""CPU, gloo, world_size 2: host-side logic of the multi-GPU paths (row-sharded k-means, single flat
gradient all-reduce). The device kernels are replaced by the oracle's CPU E/M steps (checker code)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.kmeans_oracle import assign_oracle, kmeans_oracle, make_mixture


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _CpuState:
    """stand-in for clustering.KMeansState with the same assign/accumulate/finalize protocol"""

    def __init__(self, x, K):
        self.x, self.K = x, K
        self.sums = torch.zeros(K, x.shape[1] + 1)

    def assign(self, c):
        self.labels = assign_oracle(self.x, c)
        return self.labels

    def accumulate(self):
        D = self.x.shape[1]
        self.sums.zero_()
        self.sums[:, :D].scatter_add_(0, self.labels[:, None].repeat(1, D), self.x)
        self.sums[:, D] = torch.bincount(self.labels, minlength=self.K).float()
        return self.sums

    def finalize(self, c):
        D = self.x.shape[1]
        c.copy_(self.sums[:, :D] / self.sums[:, D:D + 1])


def _kmeans_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from u2seg_b200.clustering import init_centroids_sharded, lloyd_loop
    N, D, K, Niter, seed = 2000, 32, 12, 4, 3
    x = make_mixture(N, D, 20, seed=7, spread=1.0).float()
    n_loc = N // world
    xl = x[rank * n_loc:(rank + 1) * n_loc]
    torch.manual_seed(seed)
    r = torch.randperm(N)[:K]
    c = init_centroids_sharded(xl, r, rank * n_loc, dist.group.WORLD)
    st = _CpuState(xl, K)
    lloyd_loop(st, c, Niter, dist.group.WORLD)
    gathered = [torch.empty_like(st.labels) for _ in range(world)]
    dist.all_gather(gathered, st.labels)
    if rank == 0:
        q.put((torch.cat(gathered), c.clone()))
    dist.destroy_process_group()


def test_row_sharded_kmeans_equals_single_process_oracle():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_kmeans_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    labels, c = q.get(timeout=120)
    [p.join(30) for p in procs]
    x = make_mixture(2000, 32, 20, seed=7, spread=1.0).float()
    want_l, want_c = kmeans_oracle(x, 3, K=12, Niter=4)
    assert torch.equal(labels, want_l)
    assert torch.allclose(c, want_c, rtol=1e-5, atol=1e-6)


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from u2seg_b200.engine import FlatGradients
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.Flatten(), torch.nn.Linear(4 * 36, 5))
    fg = FlatGradients(model.parameters(), torch.device("cpu"))
    g = torch.Generator().manual_seed(10 + rank)
    x = torch.randn(2, 3, 8, 8, generator=g)
    fg.zero_()
    model(x).square().sum().backward()
    assert all(p.grad.data_ptr() >= fg.flat.data_ptr() for p in model.parameters())   # still views of the flat buffer
    fg.all_reduce_mean()
    if rank == 0:   # gradients in parameter order (each starts on a 64-element boundary of the flat buffer)
        assert fg.flat.numel() % 64 == 0 and all((p.grad.data_ptr() - fg.flat.data_ptr()) % 256 == 0 for p in model.parameters())
        q.put(torch.cat([p.grad.flatten() for p in model.parameters()]).clone())
    dist.destroy_process_group()


def test_flat_gradient_allreduce_equals_mean_of_rank_gradients():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    flat = q.get(timeout=120)
    [p.join(30) for p in procs]
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.Flatten(), torch.nn.Linear(4 * 36, 5))
    tot = None
    for rank in range(2):
        x = torch.randn(2, 3, 8, 8, generator=torch.Generator().manual_seed(10 + rank))
        model.zero_grad()
        model(x).square().sum().backward()
        gr = torch.cat([p.grad.flatten() for p in model.parameters()])
        tot = gr if tot is None else tot + gr
    assert torch.allclose(flat, tot / 2, rtol=1e-5, atol=1e-6)


def _trainer_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from u2seg_b200.config import get_u2seg_cfg
    from u2seg_b200.engine import Trainer
    cfg = get_u2seg_cfg(800)
    cfg.defrost()
    cfg.MODEL.DEVICE = "cpu"
    torch.manual_seed(rank)                      # ranks start from DIFFERENT weights: the broadcast must fix that
    tr = Trainer(cfg, amp_dtype=torch.bfloat16, device=torch.device("cpu"), static_graph=True)
    before = tr._master_all.double().abs().sum()
    tr.broadcast_parameters(0)
    ok_w16 = torch.equal(tr._w16_flat, tr._master_flat.bfloat16())
    views_ok = all(torch.equal(p.detach(), tr._masters[id(p)].bfloat16()) for p in tr._low_params)
    # the step's gradient exchange: every rank fills its flat buffer, one all-reduce, mean
    g = torch.Generator().manual_seed(100 + rank)
    tr.grads.flat.copy_(torch.randn(tr.grads.flat.shape, generator=g))
    mine = tr.grads.flat[:4096].clone()
    tr.grads.all_reduce_mean()
    q.put((rank, float(before), float(tr._master_all.double().abs().sum()), float(tr._master_all.double().sum()),
           ok_w16, views_ok, mine, tr.grads.flat[:4096].clone()))
    dist.destroy_process_group()


def test_static_trainer_broadcast_and_flat_allreduce_two_ranks():
    """engine.Trainer under a 2-rank job (gloo, CPU): DDP's initial broadcast (engine/defaults.py:60-79) leaves both ranks
    with rank 0's fp32 masters and refreshed bf16 compute copies; the flat gradient buffer is averaged by ONE all-reduce."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_trainer_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    out = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    [p.join(60) for p in procs]
    (_, b0, a0, s0, w0, v0, g0, m0), (_, b1, a1, s1, w1, v1, g1, m1) = out
    assert b0 != b1                                   # different initial weights ...
    assert a0 == a1 == b0 and s0 == s1                # ... identical (rank 0's) after the broadcast
    assert w0 and w1 and v0 and v1
    assert torch.allclose(m0, (g0 + g1) / 2, rtol=1e-6, atol=1e-7) and torch.equal(m0, m1)
