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


def _run_ranks(target, nresults, timeout, world=2, attempts=3):
    """spawn `world` workers (rank, world, port, queue) and collect `nresults` queue items; a rendezvous that fails
    (the free port found above can be taken by another process before the workers bind it) is retried on a new port."""
    import queue as _queue
    last = None
    for _ in range(attempts):
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=target, args=(r, world, port, q)) for r in range(world)]
        [p.start() for p in procs]
        import time
        out, t0 = [], time.time()
        while len(out) < nresults and time.time() - t0 < timeout:
            try:
                out.append(q.get(timeout=2))
            except _queue.Empty:
                if any(p.exitcode not in (None, 0) for p in procs):      # a worker died (e.g. rendezvous failed)
                    break
        if len(out) == nresults:
            [p.join(60) for p in procs]
            return out
        last = [p.exitcode for p in procs]
        for p in procs:
            if p.is_alive():
                p.terminate()
            p.join(10)
    raise AssertionError("workers did not report after %d attempts (exit codes %r)" % (attempts, last))


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
    (labels, c), = _run_ranks(_kmeans_worker, 1, 120)
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
    flat, = _run_ranks(_grad_worker, 1, 120)
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
    out = sorted(_run_ranks(_trainer_worker, 2, 300), key=lambda t: t[0])
    (_, b0, a0, s0, w0, v0, g0, m0), (_, b1, a1, s1, w1, v1, g1, m1) = out
    assert b0 != b1                                   # different initial weights ...
    assert a0 == a1 == b0 and s0 == s1                # ... identical (rank 0's) after the broadcast
    assert w0 and w1 and v0 and v1
    assert torch.allclose(m0, (g0 + g1) / 2, rtol=1e-6, atol=1e-7) and torch.equal(m0, m1)


def _overlap_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), U2B_BUCKET_MB="24")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from u2seg_b200.config import get_u2seg_cfg
    from u2seg_b200.engine import Trainer
    cfg = get_u2seg_cfg(800)
    cfg.defrost()
    cfg.MODEL.DEVICE = "cpu"
    torch.manual_seed(0)
    tr = Trainer(cfg, amp_dtype=torch.bfloat16, device=torch.device("cpu"), static_graph=True)
    tr._setup_overlap()
    ranges = list(tr._ov["range"])
    for p in tr.params:
        p.grad = None
    g = torch.Generator().manual_seed(500 + rank)
    # a backward pass whose gradient for every parameter is a known random tensor (d/dp sum(p * r) = r); a few parameters are
    # left out of the graph on purpose: their buckets must still be reduced (zero gradient) by _finish_overlap
    left_out = {3, len(tr.params) // 2, len(tr.params) - 1}
    rs = [torch.randn(p.shape, generator=g) for p in tr.params]
    tr._arm_overlap()
    sum((p.float() * r).sum() for i, (p, r) in enumerate(zip(tr.params, rs)) if i not in left_out).backward()
    fired = sum(1 for d in tr._ov["done"] if d)
    tr._finish_overlap()
    overlap = tr.grads.flat.clone()
    tr._gather_grads()                      # the same local gradients through the single-all-reduce path
    tr.grads.all_reduce_mean()
    single = tr.grads.flat.clone()
    dst = tr._upd_grads
    lo_views = [dst[i].clone() for i in sorted(left_out)]
    q.put((rank, len(ranges), ranges[0][0], ranges[-1][1], tr.grads.flat.numel(), fired, torch.equal(overlap, single),
           float(overlap.abs().max()), [float(v.abs().max()) for v in lo_views], overlap[:2048].clone()))
    dist.destroy_process_group()


def test_overlapped_bucketed_allreduce_equals_single_allreduce_two_ranks():
    """engine.Trainer's data-parallel gradient path on 2 gloo ranks: buckets of whole parameters tile the flat buffer; the
    post-accumulate-grad hooks reduce a bucket as soon as its last gradient exists, _finish_overlap reduces the buckets that
    contain parameters without a gradient (zero-filled); the result is identical to gathering everything and issuing one
    all-reduce, and identical on both ranks."""
    out = sorted(_run_ranks(_overlap_worker, 2, 400), key=lambda t: t[0])
    for rank, nb, lo0, hi_last, n, fired, same, mx, left, head in out:
        assert nb >= 8 and lo0 == 0 and hi_last == n
        assert 0 < fired < nb            # some buckets went out during backward, the ones with left-out parameters afterwards
        assert same and mx > 0
        assert all(v == 0.0 for v in left)
    assert torch.equal(out[0][-1], out[1][-1])
