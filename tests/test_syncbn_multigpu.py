"""2-GPU test (run under torchrun by tools/run_multigpu_tests.sh; skipped in the single-process pytest runs):
fused SyncBN with the in-kernel NVLink peer exchange == NCCL all-reduce path == single-process BN over the
concatenated batch."""
import os

import pytest
import torch
import torch.distributed as dist

pytestmark = pytest.mark.gpu

_MULTI = int(os.environ.get("WORLD_SIZE", "1")) > 1


@pytest.mark.skipif(not _MULTI, reason="needs torchrun with WORLD_SIZE > 1")
def test_syncbn_peer_exchange_matches_reference():
    import torch.nn.functional as F
    from u2seg_b200.modeling import fused_bn
    from u2seg_b200.modeling.backbone import SyncBatchNorm
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    for it, (C, HW) in enumerate([(64, (24, 40)), (256, (16, 16)), (2048, (4, 4)), (512, (8, 12))] * 3):
        g = torch.Generator().manual_seed(100 + it)
        xs = [torch.randn(2, C, *HW, generator=g) * 2 + 0.3 for _ in range(world)]
        gys = [torch.randn(2, C, *HW, generator=g) for _ in range(world)]
        w, b = 1 + 0.2 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
        # reference: one process, whole batch
        xr = torch.cat(xs).cuda().requires_grad_(True)
        wr, br = w.clone().cuda().requires_grad_(True), b.clone().cuda().requires_grad_(True)
        rm, rv = torch.zeros(C).cuda(), torch.ones(C).cuda()
        yr = F.relu(F.batch_norm(xr, rm, rv, wr, br, True, 0.1, 1e-5))
        yr.backward(torch.cat(gys).cuda())
        outs = {}
        dist.barrier()                # the reference pass above loads library kernels lazily: ranks drift apart by seconds
        for mode in ("1", "0"):       # peer exchange, then NCCL
            os.environ["U2B_SYNCBN_XCHG"] = mode
            bn = SyncBatchNorm(C).cuda().train()
            with torch.no_grad():
                bn.weight.copy_(w)
                bn.bias.copy_(b)
            x = xs[rank].cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
            y = fused_bn.bn_act(x, bn, None, True)
            y.backward(gys[rank].cuda().contiguous(memory_format=torch.channels_last))
            outs[mode] = (y.detach(), x.grad.clone(), bn.weight.grad.clone(), bn.running_var.clone())
            sl = slice(2 * rank, 2 * rank + 2)
            assert torch.allclose(y, yr[sl], rtol=1e-4, atol=1e-4)
            assert torch.allclose(x.grad, xr.grad[sl], rtol=1e-3, atol=1e-5)
            assert torch.allclose(bn.running_mean, rm, rtol=1e-4, atol=1e-5) and torch.allclose(bn.running_var, rv, rtol=1e-4, atol=1e-5)
            gw = bn.weight.grad.clone()
            dist.all_reduce(gw)
            assert torch.allclose(gw, wr.grad, rtol=1e-3, atol=1e-4)
        for a, c in zip(outs["1"], outs["0"]):
            assert torch.allclose(a, c, rtol=1e-5, atol=1e-6)
    os.environ["U2B_SYNCBN_XCHG"] = "1"
    dist.barrier()


@pytest.mark.skipif(not _MULTI, reason="needs torchrun with WORLD_SIZE > 1")
def test_static_graph_trainer_ranks_stay_in_sync():
    """Data-parallel static-graph step (engine.Trainer: SyncBN peer exchange + one flat NCCL all-reduce + fused
    optimizer, all inside the captured graph): ranks see different batches, start from identical weights, and must
    hold bit-identical fp32 masters afterwards (DDP invariant, engine/defaults.py:60-79); losses finite."""
    from u2seg_b200.config import get_u2seg_cfg
    from u2seg_b200.data_synth import synthetic_batch
    from u2seg_b200.engine import Trainer
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    torch.manual_seed(0)
    tr = Trainer(get_u2seg_cfg(800), amp_dtype=torch.bfloat16, static_graph=True)
    tr.broadcast_parameters(0)
    batch = synthetic_batch(2, 256, 320, 800, 28, seed=40 + rank, G=6, min_size=24, max_size=160)
    torch.manual_seed(100 + rank)
    hist = [float(sum(tr.run_step(batch).values())) for _ in range(3)]
    torch.cuda.synchronize()
    assert all(h == h and h < 1e4 for h in hist), hist
    digest = torch.stack([tr._master_all.double().sum(), tr._master_all.double().abs().sum(),
                          tr._mom_all.double().abs().sum()])
    gathered = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(gathered, digest)
    for g in gathered[1:]:
        assert torch.equal(g, gathered[0]), (gathered[0].tolist(), g.tolist())
    losses = torch.tensor(hist, device="cuda", dtype=torch.float64)
    other = [torch.zeros_like(losses) for _ in range(world)]
    dist.all_gather(other, losses)
    assert not torch.equal(other[0], other[1])          # ranks really worked on different data
    dist.barrier()


@pytest.mark.skipif(not _MULTI, reason="needs torchrun with WORLD_SIZE > 1")
def test_overlapped_bucketed_allreduce_equals_single_allreduce(monkeypatch):
    """engine.Trainer: gradient buckets all-reduced on a side stream while backward runs (post-accumulate-grad hooks) give
    the same averaged flat gradient buffer as the single all-reduce issued after backward."""
    from u2seg_b200.config import get_u2seg_cfg
    from u2seg_b200.data_synth import synthetic_batch
    from u2seg_b200.engine import Trainer
    from u2seg_b200.modeling import rpn, static_train
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    if not dist.is_initialized():
        dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    monkeypatch.setattr(rpn, "_randperm", lambda n, device=None: torch.arange(n, device=device))
    monkeypatch.setattr(static_train, "_rand_keys",
                        lambda mask: torch.arange(mask.numel(), device=mask.device, dtype=torch.float32).view(mask.shape) / (mask.numel() + 1))
    batch = synthetic_batch(2, 256, 320, 800, 28, seed=70 + rank, G=6, min_size=24, max_size=160)
    flats, ranges = {}, None
    # Two runs of the SAME configuration are not bit-identical (fp32 atomics in the ROI-align / upsampling backward, then bf16
    # roundings downstream): the single-all-reduce path runs twice and its own run-to-run difference is the yardstick.
    for name, overlap in (("overlap", "1"), ("single", "0"), ("single_again", "0")):
        monkeypatch.setenv("U2B_OVERLAP_ALLREDUCE", overlap)
        cfg = get_u2seg_cfg(800)
        cfg.SOLVER.BASE_LR = 0.0
        torch.manual_seed(0)
        tr = Trainer(cfg, amp_dtype=torch.bfloat16, static_graph=True)
        tr.broadcast_parameters(0)
        tr.run_step(batch)
        torch.cuda.synchronize()
        flats[name] = tr.grads.flat.clone()
        assert (overlap == "1") == (getattr(tr, "_ov", None) is not None)
        if overlap == "1":
            ranges = list(tr._ov["range"])
            assert len(ranges) >= 4 and ranges[0][0] == 0 and ranges[-1][1] == tr.grads.flat.numel()
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))            # buckets tile the whole buffer
        del tr
    a, b, c = flats["overlap"], flats["single"], flats["single_again"]
    assert float(a.abs().max()) > 0

    def rel(u, v):
        return float((u - v).double().norm() / (v.double().norm() + 1e-30))

    print("overlap vs single %.2e; single vs single (run-to-run) %.2e" % (rel(a, b), rel(c, b)))
    assert rel(a, b) <= 3 * rel(c, b) + 1e-6
    for lo, hi in ranges:                # a bucket reduced too early / twice / never would stand out on its own
        assert rel(a[lo:hi], b[lo:hi]) <= 5 * rel(c[lo:hi], b[lo:hi]) + 1e-5, (lo, hi, rel(a[lo:hi], b[lo:hi]), rel(c[lo:hi], b[lo:hi]))
    ga = [torch.zeros_like(a[:1000]) for _ in range(world)]
    dist.all_gather(ga, a[-1000:].contiguous())
    assert torch.equal(ga[0], ga[1])                                    # every rank holds the same averaged gradients
    dist.barrier()
