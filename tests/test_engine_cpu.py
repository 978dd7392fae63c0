"""Host logic of the static-graph trainer (u2seg_b200/engine.py) on CPU: flat master / gradient / momentum buffers,
bf16 compute views, segment tables of the fused optimizer kernel, gradient gather, and the foreach optimizer path
against torch.optim.SGD + per-parameter clip_grad_norm_ (detectron2/solver/build.py:63-73,119-139). No kernels run."""
import pytest
import torch


@pytest.fixture(scope="module")
def trainer():
    from u2seg_b200.config import get_u2seg_cfg
    from u2seg_b200.engine import Trainer
    cfg = get_u2seg_cfg(800)
    cfg.defrost()
    cfg.MODEL.DEVICE = "cpu"
    torch.manual_seed(0)
    return Trainer(cfg, amp_dtype=torch.bfloat16, device=torch.device("cpu"), static_graph=True)


def _span(t):
    return t.data_ptr(), t.data_ptr() + t.numel() * t.element_size()


def test_flat_buffers_layout(trainer):
    tr = trainer
    assert tr.lowp and len(tr.params) == 248
    lo, hi = _span(tr._master_all)
    spans = sorted(_span(m) for m in tr._upd_params)
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:])), "master views overlap"
    assert all(lo <= a and b <= hi for a, b in spans)
    assert all((a - lo) % 256 == 0 for a, _ in spans)                       # 64-element (256-byte) boundaries
    glo, ghi = _span(tr.grads.flat)
    assert tr.grads.flat.numel() == tr._master_all.numel() == tr._mom_all.numel()
    for m, g in zip(tr._upd_params, tr._upd_grads):                         # same offsets in master and gradient buffers
        assert m.data_ptr() - lo == g.data_ptr() - glo and m.shape == g.shape and m.stride() == g.stride()
        assert m.dtype == torch.float32 and g.dtype == torch.float32
    wlo, whi = _span(tr._w16_flat)
    n_head = tr.grads.n_head
    for p in tr._low_params:                                                # bf16 compute copies mirror the master tail
        m = tr._masters[id(p)]
        assert p.dtype == torch.bfloat16 and wlo <= p.data_ptr() < whi
        assert (p.data_ptr() - wlo) // 2 == (m.data_ptr() - lo) // 4 - n_head
        assert torch.equal(p.detach(), m.bfloat16())
    low = {id(p) for p in tr._low_params}
    for p in tr.params:
        if id(p) not in low:                                                # norm parameters ARE views of the master buffer
            assert p.dtype == torch.float32 and lo <= p.data_ptr() < lo + n_head * 4


def test_segment_tables_and_names(trainer):
    tr = trainer
    sv = tr.cfg.SOLVER
    seg = tr._seg_chunk
    assert seg.numel() * 64 == tr._master_all.numel() and int(seg.min()) == 0 and int(seg.max()) == 247
    assert bool((seg[1:] >= seg[:-1]).all())                                # buffer order
    counts = torch.bincount(seg.long(), minlength=248)
    lo = tr._master_all.data_ptr()
    seg_params = [p for p in tr.params if id(p) not in {id(q) for q in tr._low_params}] + tr._low_params
    by_id = dict(zip((id(p) for p in tr.params), tr._upd_params))
    names = dict((id(p), n) for n, p in tr.model.named_parameters())
    for i, p in enumerate(seg_params):
        m = by_id[id(p)]
        first = (m.data_ptr() - lo) // 4 // 64
        assert int(seg[first]) == i and int(counts[i]) == (p.numel() + 63) // 64
        is_norm = ".norm." in names[id(p)]
        assert float(tr._seg_wd[i]) == pytest.approx(sv.WEIGHT_DECAY_NORM if is_norm else sv.WEIGHT_DECAY)
        assert tr._seg_grads[i].data_ptr() - tr.grads.flat.data_ptr() == m.data_ptr() - lo
    mp = tr.master_parameters()
    want = {n: p.shape for n, p in tr.model.named_parameters() if p.requires_grad}
    assert list(mp) == list(want) and all(mp[n].shape == want[n] and mp[n].dtype == torch.float32 for n in want)


def test_gather_then_foreach_optimizer_equals_torch_sgd(trainer):
    tr = trainer
    sv = tr.cfg.SOLVER
    g = torch.Generator().manual_seed(7)
    ref = [m.detach().clone().requires_grad_(True) for m in tr._upd_params]
    norm = {id(p) for n, p in tr.model.named_parameters() if ".norm." in n}
    opt = torch.optim.SGD([{"params": [r], "weight_decay": sv.WEIGHT_DECAY_NORM if id(p) in norm else sv.WEIGHT_DECAY}
                           for r, p in zip(ref, tr.params)], lr=0.02, momentum=sv.MOMENTUM, nesterov=sv.NESTEROV)
    tr._lr_t.fill_(0.02)
    for step, scale in enumerate((1e-3, 5.0)):                              # the second step clips
        for p, r in zip(tr.params, ref):
            gr = torch.randn(p.shape, generator=g) * scale
            p.grad = gr.to(p.dtype).contiguous(memory_format=torch.channels_last) if p.dim() == 4 else gr.to(p.dtype)
            r.grad = p.grad.float().clone()                                 # what the fp32 buffer must receive
        tr._gather_grads()
        for dst, r in zip(tr._upd_grads, ref):
            assert torch.equal(dst, r.grad)
        for r in ref:
            torch.nn.utils.clip_grad_norm_(r, tr.clip.CLIP_VALUE, tr.clip.NORM_TYPE)
        opt.step()
        tr._clip_foreach()
        tr._sgd_foreach()
        for m, r in zip(tr._upd_params, ref):
            assert torch.allclose(m, r.detach(), rtol=1e-5, atol=1e-7), step
        assert torch.equal(tr._w16_flat, tr._master_flat.bfloat16())
    for p in tr.params:
        p.grad = None


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_decode_topk_level_equals_decode_all_then_gather(dtype):
    """static_train.decode_topk_level (decode only the top-k anchors) is BIT-identical to the reference order
    (rpn.py:497-533 decode every anchor, proposal_utils.py:67-84 top-k, gather)."""
    from u2seg_b200.modeling.rpn import Box2BoxTransform
    from u2seg_b200.modeling.static_train import decode_topk_level
    g = torch.Generator().manual_seed(11)
    N, A, k = 2, 5000, 300
    b2b = Box2BoxTransform(weights=(1.0, 1.0, 1.0, 1.0))
    c = torch.rand(A, 2, generator=g) * 512
    wh = torch.rand(A, 2, generator=g) * 200 + 4
    anchors = torch.cat([c - wh / 2, c + wh / 2], 1)
    logits = torch.randn(N, A, generator=g).to(dtype)
    deltas = (torch.randn(N, A, 4, generator=g) * 0.5).to(dtype)
    sc, boxes = decode_topk_level(b2b, anchors, logits, deltas, k)
    props = b2b.apply_deltas(deltas.reshape(-1, 4), anchors.unsqueeze(0).expand(N, -1, -1).reshape(-1, 4)).view(N, -1, 4)
    sc_ref, idx = logits.float().topk(k, dim=1)
    want = props[torch.arange(N)[:, None], idx]
    assert torch.equal(sc, sc_ref) and torch.equal(boxes, want) and boxes.dtype == torch.float32


def test_subsample_static_matches_reference_sampling_semantics(monkeypatch):
    """static_train.subsample_static (random-key top-k over fixed shapes) against sampling.py:9-54 subsample_labels:
    same counts (num_pos = min(#pos, int(n*frac)), num_neg = min(#neg, n - num_pos)), indices drawn from the right
    candidate sets without repetition, positives first; and with the key generator replaced by the inverse of a fixed
    permutation it selects EXACTLY `candidates[perm][:k]` (INT: bit exact)."""
    from u2seg_b200.modeling import static_train
    g = torch.Generator().manual_seed(3)
    for n_pos, n_neg, n, frac in [(10, 5000, 256, 0.5), (300, 40, 256, 0.5), (0, 100, 64, 0.25), (700, 900, 512, 0.25)]:
        total = n_pos + n_neg + 37
        perm = torch.randperm(total, generator=g)
        is_pos = torch.zeros(total, dtype=torch.bool)
        is_neg = torch.zeros(total, dtype=torch.bool)
        is_pos[perm[:n_pos]] = True
        is_neg[perm[n_pos:n_pos + n_neg]] = True
        idx, ok, fg = static_train.subsample_static(is_pos, is_neg, n, frac)
        want_pos = min(n_pos, int(n * frac))
        want_neg = min(n_neg, n - want_pos)
        assert idx.shape == ok.shape == fg.shape == (n,)
        assert int(fg.sum()) == want_pos and int(ok.sum()) == want_pos + want_neg
        assert bool(ok[:want_pos + want_neg].all()) and not bool(ok[want_pos + want_neg:].any())     # real slots first
        assert bool(fg[:want_pos].all())                                                             # positives first
        sel = idx[ok]
        assert sel.unique().numel() == sel.numel()                                                   # no repetition
        assert bool(is_pos[idx[fg]].all()) and bool(is_neg[idx[ok & ~fg]].all())
    # deterministic keys: rank of each element in a fixed permutation -> the reference's `candidates[randperm][:k]`
    total, k = 1000, 64
    mask = torch.rand(total, generator=g) < 0.3
    order = torch.randperm(total, generator=g)
    rank = torch.empty(total)
    rank[order] = torch.arange(total, dtype=torch.float32) / total
    monkeypatch.setattr(static_train, "_rand_keys", lambda m: rank.clone())
    idx, ok = static_train._topk_select(mask, k)
    want = order[mask[order]][:k]
    assert torch.equal(idx[ok], want)


def test_pack_batch_pads_to_fixed_capacity():
    """static_train.pack_batch: reference-format list[dict] -> fixed-capacity tensors (g_max GT slots with a validity
    mask; images NHWC bytes). Padding slots must be inert: zero boxes, invalid, empty masks."""
    from u2seg_b200.data_synth import synthetic_batch
    from u2seg_b200.modeling.static_train import pack_batch
    batch = synthetic_batch(2, 96, 128, 800, 28, seed=9, G=5, min_size=12, max_size=60, device="cpu")
    imgs, gb, gc, gv, gm, sem = pack_batch(batch, torch.device("cpu"), g_max=8)
    assert imgs.shape == (2, 3, 96, 128) and imgs.dtype == torch.uint8 and imgs.stride(1) == 1         # NHWC storage
    assert gb.shape == (2, 8, 4) and gc.shape == (2, 8) and gv.shape == (2, 8) and gm.shape == (2, 8, 96, 128)
    assert sem.shape == (2, 96, 128)
    for n, d in enumerate(batch):
        inst = d["instances"]
        g = len(inst)
        assert torch.equal(imgs[n], d["image"]) and torch.equal(sem[n], d["sem_seg"])
        assert torch.equal(gb[n, :g], inst.gt_boxes.tensor) and torch.equal(gc[n, :g], inst.gt_classes)
        assert torch.equal(gm[n, :g], inst.gt_masks.tensor)
        assert bool(gv[n, :g].all()) and not bool(gv[n, g:].any())
        assert float(gb[n, g:].abs().sum()) == 0 and not bool(gm[n, g:].any())


def test_trainer_state_dict_round_trip(trainer):
    """Trainer.state_dict(): reference names / shapes, fp32 values of the masters (not the bf16 compute copies);
    load_state_dict() restores masters, bf16 copies and buffers."""
    tr = trainer
    sd = tr.state_dict()
    ref = tr.model.state_dict()
    assert list(sd) == list(ref) and len(sd) == 431
    pnames = {n for n, _ in tr.model.named_parameters()}
    assert all(sd[k].shape == ref[k].shape for k in sd)
    assert all(sd[k].dtype == torch.float32 for k in pnames)
    low = next(n for n, p in tr.model.named_parameters() if p.dtype == torch.bfloat16)
    assert not torch.equal(sd[low], ref[low].float())                  # masters carry more bits than the bf16 copies
    assert torch.equal(sd[low].bfloat16(), ref[low])
    sd2 = {k: (v * 0.5 if v.is_floating_point() else v) for k, v in sd.items()}
    tr.load_state_dict(sd2)
    for k, v in tr.state_dict().items():
        assert torch.equal(v, sd2[k]), k
    assert torch.equal(tr._w16_flat, tr._master_flat.bfloat16())
    tr.load_state_dict(sd)


def test_trainer_checkpoint_round_trip_and_non_strict_load(trainer):
    """Trainer.checkpoint(): model + momentum + iteration (what the reference checkpointer saves); load_state_dict
    strict=False returns (missing, unexpected) like DetectionCheckpointer's backbone-only pretrain load; shape
    mismatches raise instead of broadcasting."""
    tr = trainer
    ck = tr.checkpoint()
    assert set(ck) == {"model", "optimizer", "iteration", "scaler"} and len(ck["optimizer"]["momentum"]) == 248
    mom = tr._momentum_by_name()
    name = "backbone.fpn_output3.weight"
    assert mom[name].shape == tr.master_parameters()[name].shape
    mom[name].fill_(0.25)
    assert float(tr._mom_all.sum()) == pytest.approx(0.25 * mom[name].numel())   # a view of the flat momentum buffer
    ck2 = tr.checkpoint()
    mom[name].zero_()
    it = tr.iter
    tr.iter = 123
    tr.load_checkpoint(ck2)
    assert float(tr._momentum_by_name()[name].mean()) == 0.25 and tr.iter == it
    tr._mom_all.zero_()
    # backbone-only pretrain: every other key is reported missing, nothing raises
    backbone_only = {k: v for k, v in ck["model"].items() if k.startswith("backbone.bottom_up.")}
    backbone_only["fc1000.weight"] = torch.zeros(3)
    missing, unexpected = tr.load_state_dict(backbone_only, strict=False)
    assert unexpected == ["fc1000.weight"] and all(not k.startswith("backbone.bottom_up.") for k in missing) and missing
    with pytest.raises(AssertionError):
        tr.load_state_dict(backbone_only)
    bad = dict(ck["model"])
    bad[name] = bad[name][:1]
    with pytest.raises(ValueError):
        tr.load_state_dict(bad)
    tr.load_state_dict(ck["model"])
