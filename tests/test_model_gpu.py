"""GPU parity of the whole detector step (product u2seg_b200 model, fp32, channels_last) against the
reference's recorded outputs (tests/golden/detector_{train,infer}_*.npz) and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import detector_oracle as do

pytestmark = pytest.mark.gpu


def _cpu_randperm(n, device=None):
    return torch.randperm(n).to(device)     # CPU generator, the order the oracle/reference consume it in


def _make_batch(data, train=True, out_sizes=None):
    from u2seg_b200.structures import BitMasks, Boxes, Instances
    images, boxes, classes, masks, sems = data
    batch = []
    for i, im in enumerate(images):
        d = {"image": im}
        if train:
            inst = Instances((im.shape[1], im.shape[2]))
            inst.gt_boxes = Boxes(boxes[i])
            inst.gt_classes = classes[i]
            inst.gt_masks = BitMasks(masks[i])
            d["instances"] = inst
            d["sem_seg"] = sems[i]
        elif out_sizes is not None:
            d["height"], d["width"] = out_sizes[i]
        batch.append(d)
    return batch


def _build(K, params, training):
    from u2seg_b200.config import get_u2seg_cfg
    from u2seg_b200.modeling import build_model
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = get_u2seg_cfg(K)
    model = build_model(cfg)
    model.load_state_dict(params)
    model = model.to(memory_format=torch.channels_last)
    model.train(training)
    return model


def test_state_dict_names_and_shapes():
    model = _build(800, do.init_params(do.DetCfg(800), 0), True)
    sd = model.state_dict()
    p = do.init_params(do.DetCfg(800), 0)
    assert set(sd) == set(p) and len(sd) == 431 and len(list(model.parameters())) == 248


def test_training_losses_match_reference(golden_dir, monkeypatch):
    from u2seg_b200.modeling import rpn
    g = np.load(os.path.join(golden_dir, "detector_train_256x320.npz"))
    n, H, W, K, S, seed, G, lo, hi = [int(v) for v in g["meta"]]
    model = _build(K, do.init_params(do.DetCfg(K, S), 0), True)
    data = do.synthetic_batch(n, H, W, K, S, seed=seed, G=G, min_size=lo, max_size=hi)
    monkeypatch.setattr(rpn, "_randperm", _cpu_randperm)
    torch.manual_seed(seed)
    losses = model(_make_batch(data))
    assert list(losses.keys()) == [str(k) for k in g["keys"]]
    for k, v in zip(g["keys"], g["values"]):
        got = float(losses[str(k)])
        assert abs(got - v) <= 1e-3 * max(1.0, abs(v)), (str(k), got, v)     # FP: within 1e-3 (fp32)
    sum(losses.values()).backward()
    grads = [p.grad for p in model.parameters()]
    assert all(gr is not None and torch.isfinite(gr).all() for gr in grads)


def test_training_gradients_match_oracle(monkeypatch):
    """Backward parity: d(sum of losses)/d(params) vs autograd through the CPU oracle, 2 images of 256x256 (BN batches are not
    degenerate). Both sides sample "the first k candidates in index order" (randperm -> arange, as in the BASELINE-size
    fixtures): with random draws one borderline candidate changes the length of a permutation and with it every later
    sample, which shows up as a few per cent in the gradients and says nothing about the arithmetic."""
    from u2seg_b200.modeling import rpn
    K, S, seed = 800, 28, 5
    cfg = do.DetCfg(K, S)
    params = do.init_params(cfg, 0)
    data = do.synthetic_batch(2, 256, 256, K, S, seed=seed, G=6, min_size=16, max_size=120)
    names = ["backbone.fpn_output2.weight", "backbone.bottom_up.res3.0.conv2.weight", "roi_heads.box_head.1.fc1.weight",
             "roi_heads.mask_head.mask_fcn2.weight", "proposal_generator.rpn_head.conv.weight", "sem_seg_head.p4.2.weight",
             "backbone.bottom_up.stem.conv1.norm.weight"]
    op = {k: v.clone().requires_grad_(k in names) for k, v in params.items()}

    def first(n, device=None, **kw):
        return torch.arange(n, device=device)

    with monkeypatch.context() as m:
        m.setattr(torch, "randperm", first)
        want_losses = do.forward_train(op, cfg, *data)
        sum(want_losses.values()).backward()
    model = _build(K, params, True)
    monkeypatch.setattr(rpn, "_randperm", first)
    got_losses = model(_make_batch(data))
    sum(got_losses.values()).backward()
    worst = max(abs(float(got_losses[k]) - float(want_losses[k])) / max(abs(float(want_losses[k])), 1e-3) for k in want_losses)
    print("   worst loss error %.2e" % worst)
    assert worst < 1e-4
    named = dict(model.named_parameters())
    bad = []
    for k in names:
        a, b = named[k].grad.double().cpu(), op[k].grad.double()
        l2 = float((a - b).norm() / (b.norm() + 1e-30))
        mx = float((a - b).abs().max() / (b.abs().max() + 1e-30))
        print("   %-45s relative L2 error %.2e, max error / max entry %.2e" % (k, l2, mx))
        # fp32 on both sides, identical discrete decisions (losses agree to 5e-6). Heads / FPN / RPN: relative L2 error
        # 7e-4 .. 3e-3 measured, bound 1e-2. Backbone parameters sit below up to 53 batch-norm layers that amplify fp32
        # summation-order noise: 2.0e-2 (res3.0.conv2) and 1.8e-2 (stem norm) measured - the size of the reference's OWN
        # fp32-vs-fp64 deviation at the BASELINE size (2.1e-2 on res4.0.conv1.weight; tests/test_baseline_config_gpu.py uses
        # the float64 yardstick to make that comparison tight). A wrong term in a backward formula shows up as O(1).
        deep = k.startswith("backbone.bottom_up")
        if l2 > (5e-2 if deep else 1e-2) or mx > (5e-2 if deep else 3e-2):
            bad.append((k, l2, mx))
    assert not bad, bad


def test_inference_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "detector_infer_200x304.npz"))
    n, H, W, K, S, seed, G, lo, hi, oh, ow = [int(v) for v in g["meta"]]
    cfg = do.DetCfg(K, S)
    data = do.synthetic_batch(n, H, W, K, S, seed=seed, G=G, min_size=lo, max_size=hi)
    model = _build(K, do.eval_fixture_params(cfg, data[0], seed=0), False)
    out = model(_make_batch(data, train=False, out_sizes=[(oh, ow)]))[0]
    inst = out["instances"]
    # detections are ordered by score; fp32 rounding may permute near-equal scores, so match each reference
    # detection to ours by (class, box) and require identical classes + boxes/scores within 1e-3
    gb, gs, gc = inst.pred_boxes.tensor.cpu().numpy(), inst.scores.cpu().numpy(), inst.pred_classes.cpu().numpy()
    assert len(gc) == len(g["pred_classes"])
    np.testing.assert_allclose(np.sort(gs)[::-1], g["scores"], rtol=1e-3, atol=1e-5)
    want_masks = np.unpackbits(g["pred_masks"])[:int(np.prod(g["mask_shape"]))].reshape(g["mask_shape"]).astype(bool)
    got_masks = inst.pred_masks.cpu().numpy()
    used = set()
    for j in range(len(gc)):
        cand = [i for i in range(len(gc)) if i not in used and gc[i] == g["pred_classes"][j]
                and np.abs(gb[i] - g["pred_boxes"][j]).max() <= 1e-2 + 1e-3 * np.abs(g["pred_boxes"][j]).max()]
        if not cand:        # the tail of the top-100 list (near-equal scores at the cut) may swap members
            assert j >= 90, ("no match for reference detection", j)
            continue
        i = cand[0]
        used.add(i)
        assert abs(gs[i] - g["scores"][j]) <= 1e-3 * g["scores"][j] + 1e-5
        assert (got_masks[i] != want_masks[j]).mean() < 1e-3     # FP->bool: only pixels with |p-0.5| < eps may flip
    sem = out["sem_seg"].argmax(0).cpu().numpy().astype(np.uint8)
    assert (sem != g["sem_seg_argmax"]).mean() < 1e-3
    pan, info = out["panoptic_seg"]
    assert (pan.cpu().numpy() != g["panoptic"]).mean() < 2e-3 and len(info) == int(g["n_segments"][0])


def test_bf16_step_with_tcgen05_convs_close_to_library_convs(monkeypatch):
    """autocast(bf16) training step: large 3x3 convs on conv_tc (fwd + dgrad) vs the library kernels — same
    sampling, losses within 2%, finite gradients that agree in norm."""
    from u2seg_b200.modeling import ops, rpn
    K, S, seed = 800, 28, 9
    cfg = do.DetCfg(K, S)
    params = do.init_params(cfg, 0)
    data = do.synthetic_batch(2, 256, 320, K, S, seed=seed, G=6, min_size=24, max_size=160)
    monkeypatch.setattr(rpn, "_randperm", _cpu_randperm)
    out = {}
    for policy in ("none", "all"):
        monkeypatch.setattr(ops, "TCGEN05_CONV_POLICY", policy)
        model = _build(K, params, True)
        torch.manual_seed(seed)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            losses = model(_make_batch(data))
        sum(losses.values()).backward()
        gn = {n: float(p.grad.float().norm()) for n, p in model.named_parameters()}
        out[policy] = ({k: float(v) for k, v in losses.items()}, gn)
    (la, ga), (lb, gb) = out["none"], out["all"]
    for k in la:
        # the box-regression losses average over the few foreground ROIs that survive the (discontinuous) top-k + NMS
        # proposal selection, which two bf16 implementations resolve differently: looser bound there
        tol = 5e-2 if "box_reg" in k else 2e-2
        assert abs(la[k] - lb[k]) <= tol * max(1.0, abs(la[k])), (k, la[k], lb[k])
    for n in ("backbone.fpn_output2.weight", "proposal_generator.rpn_head.conv.weight", "roi_heads.mask_head.mask_fcn1.weight",
              "sem_seg_head.p2.0.weight", "backbone.bottom_up.res2.0.conv1.weight"):
        assert all(map(lambda v: v == v and v < float("inf"), (ga[n], gb[n])))
        assert abs(ga[n] - gb[n]) <= 0.1 * max(ga[n], 1e-8), (n, ga[n], gb[n])


def test_static_path_equals_dynamic_path(monkeypatch):
    """modeling/static_train.py (fixed-capacity buffers, no host sync) computes the same losses and gradients as the
    reference-shaped dynamic path when both samplers are made deterministic ("first k in index order")."""
    from u2seg_b200.modeling import rpn, static_train
    K, S, seed = 800, 28, 13
    cfg = do.DetCfg(K, S)
    params = do.init_params(cfg, 0)
    data = do.synthetic_batch(2, 192, 256, K, S, seed=seed, G=5, min_size=20, max_size=120)
    monkeypatch.setattr(rpn, "_randperm", lambda n, device=None: torch.arange(n, device=device))
    monkeypatch.setattr(static_train, "_rand_keys",
                        lambda mask: torch.arange(mask.numel(), device=mask.device, dtype=torch.float32) / (mask.numel() + 1))
    batch = _make_batch(data)
    model = _build(K, params, True)
    dyn = model(batch)
    sum(dyn.values()).backward()
    gd = {n: p.grad.clone() for n, p in model.named_parameters()}
    model2 = _build(K, params, True)
    packed = static_train.pack_batch(batch, torch.device("cuda"), g_max=8)      # 3 padded GT slots per image
    sta, flag = static_train.forward_train_static(model2, *packed)
    assert not bool(flag)
    sum(sta.values()).backward()
    assert list(sta.keys()) == list(dyn.keys())
    for k in dyn:
        assert abs(float(sta[k]) - float(dyn[k])) <= 1e-4 * max(1.0, abs(float(dyn[k]))), (k, float(sta[k]), float(dyn[k]))
    for n, p in model2.named_parameters():
        d = float(gd[n].abs().max()) + 1e-12
        assert float((p.grad - gd[n]).abs().max()) <= 2e-3 * d + 1e-7, n


def test_static_graph_trainer_runs_and_learns():
    """Trainer(static_graph=True): whole step (fwd+bwd+clip+SGD) replayed from one CUDA graph; losses finite, identical
    inputs give a decreasing total loss, parameters change, no host sync needed between steps."""
    from u2seg_b200.config import get_u2seg_cfg
    from u2seg_b200.data_synth import synthetic_batch
    from u2seg_b200.engine import Trainer
    torch.manual_seed(0)
    cfg = get_u2seg_cfg(800)
    tr = Trainer(cfg, amp_dtype=torch.bfloat16, static_graph=True)
    batch = synthetic_batch(2, 256, 320, 800, 28, seed=3, G=6, min_size=24, max_size=160)
    w0 = tr.model.backbone.fpn_output3.weight.detach().clone()
    hist = []
    for _ in range(12):
        losses = tr.run_step(batch)
        hist.append(float(sum(losses.values())))
    tr.check_finite()
    assert all(h == h and h < 1e4 for h in hist)
    assert hist[-1] < hist[0]
    assert not torch.equal(tr.model.backbone.fpn_output3.weight.detach(), w0)


def test_static_trainer_fused_optimizer_equals_foreach_optimizer():
    """Trainer(static_graph=True) optimizer step in isolation: the fused clip+SGD+bf16-refresh kernel over the flat
    master buffer (csrc/optimizer.cu) against the foreach path (torch multi-tensor ops on the per-parameter views,
    solver/build.py:63-73,119-139) on the real parameter set, fed identical random gradients for 3 steps (the second
    one large enough to clip). Also checks that the flat-buffer views do not alias."""
    from u2seg_b200.config import get_u2seg_cfg
    from u2seg_b200.engine import Trainer
    cfg = get_u2seg_cfg(800)
    trainers = []
    for _ in range(2):
        torch.manual_seed(0)
        tr = Trainer(cfg, amp_dtype=torch.bfloat16, static_graph=True)
        views = sorted((m.data_ptr(), m.data_ptr() + m.numel() * 4) for m in tr._upd_params)
        assert all(a[1] <= b[0] for a, b in zip(views, views[1:])), "master views overlap"
        lo, hi = tr._master_all.data_ptr(), tr._master_all.data_ptr() + tr._master_all.numel() * 4
        assert all(lo <= a and b <= hi for a, b in views)
        trainers.append(tr)
    a, b = trainers
    for x, y in zip(a._upd_params, b._upd_params):
        assert torch.equal(x, y)
    g = torch.Generator(device="cuda").manual_seed(5)
    for step, scale in enumerate((1e-3, 3.0, 1e-2)):
        for tr in (a, b):
            tr._lr_t.fill_(0.01 * (step + 1))
        for ga, gb in zip(a._upd_grads, b._upd_grads):        # identical gradients in both trainers' flat buffers
            r = torch.randn(ga.shape, generator=g, device="cuda") * scale
            ga.copy_(r)
            gb.copy_(r)
        a._fused_clip_sgd()
        b._clip_foreach()
        b._sgd_foreach()
        torch.cuda.synchronize()
        for (name, ma), mb in zip(a.master_parameters().items(), b.master_parameters().values()):
            assert torch.allclose(ma, mb, rtol=1e-5, atol=1e-8), (step, name)                 # FLOAT: fma contraction only
        assert torch.equal(a._w16_flat, a._master_flat.bfloat16())                            # bf16 refresh: exact rounding
        for pa, pb in zip(a.params, b.params):                                                # what the modules see
            assert torch.allclose(pa.detach().float(), pb.detach().float(), rtol=1e-2, atol=1e-8)   # bf16 copies: 1 ulp
