"""Parity at the BASELINE.json configuration itself (config 2: u2seg_R50_800 training step, 2 x 1024x1024 synthetic
images, G=20 instances per image), against values recorded from the UNMODIFIED reference on CPU
(tests/golden/detector_train_1024_{randperm,first}.npz, produced by `python oracle/make_golden.py baseline`).

  * fp32, reference-shaped dynamic path and fixed-capacity static path (what the benchmark runs), sampler "first k
    candidates in index order" on both sides: the 10 losses within 1e-3 (measured: 5e-7), per-parameter gradient L2
    norms of all 248 parameters as close to the reference's float64 run as the reference's own float32 run is;
  * fp32 dynamic path with the reference's own CPU randperm draws injected: losses within 1e-3;
  * bf16 autocast, static path replayed from the whole-step CUDA graph (the benchmarked configuration: tcgen05 convs,
    fused bf16 BN, selected-class mask predictor): losses and gradient norms within bf16 bounds (stated in the test).

Gradient tolerance. The backward pass runs through 61 batch-norm layers; in float32 some gradient norms are reproducible
only to a few 1e-3 whatever the implementation: the reference's float32 run deviates from its own float64 run
(detector_train_1024_first_fp64.npz, `make_golden.py baseline64`) by 1.1e-4 (median), 2.1e-3 (95th percentile),
5.4e-3 (worst parameter). The tests measure the product against the float64 values and require each of these three
statistics to stay within 2x of the reference's own.
"""
import os

import numpy as np
import pytest
import torch

from oracle import detector_oracle as do

pytestmark = pytest.mark.gpu


def _cpu_randperm(n, device=None):
    return torch.randperm(n).to(device)


def _first(n, device=None):
    return torch.arange(n, device=device)


def _first_keys(mask):
    return torch.arange(mask.numel(), device=mask.device, dtype=torch.float32).view(mask.shape) / (mask.numel() + 1)


def _golden(golden_dir, sampler):
    g = np.load(os.path.join(golden_dir, "detector_train_1024_%s.npz" % sampler))
    n, H, W, K, S, seed, G, lo, hi = [int(v) for v in g["meta"]]
    data = do.synthetic_batch(n, H, W, K, S, seed=seed, G=G, min_size=lo, max_size=hi)
    return g, (n, H, W, K, S, seed, G), data


def _batch(data):
    from u2seg_b200.structures import BitMasks, Boxes, Instances
    images, boxes, classes, masks, sems = data
    out = []
    for i, im in enumerate(images):
        inst = Instances((im.shape[1], im.shape[2]))
        inst.gt_boxes = Boxes(boxes[i])
        inst.gt_classes = classes[i]
        inst.gt_masks = BitMasks(masks[i])
        out.append({"image": im, "instances": inst, "sem_seg": sems[i]})
    return out


def _build_fp32(K, S):
    from u2seg_b200.config import get_u2seg_cfg
    from u2seg_b200.modeling import build_model
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = build_model(get_u2seg_cfg(K))
    model.load_state_dict(do.init_params(do.DetCfg(K, S), 0))
    model = model.to(memory_format=torch.channels_last)
    model.train(True)
    return model


def _check_losses(losses, g, rtol):
    assert list(losses.keys()) == [str(k) for k in g["keys"]]
    worst = 0.0
    for k, v in zip(g["keys"], g["values"]):
        got = float(losses[str(k)])
        err = abs(got - v) / max(1.0, abs(v))
        worst = max(worst, err)
        assert err <= rtol, (str(k), got, float(v), err)
    return worst


def _grad_norm_errors(named_grads, g):
    want = dict(zip([str(n) for n in g["grad_names"]], g["grad_norms"]))
    assert set(want) == set(named_grads), set(want) ^ set(named_grads)
    scale = float(np.median(list(want.values())))
    errs = {}
    for n, w in want.items():
        got = float(named_grads[n].double().norm())
        errs[n] = abs(got - w) / max(w, 1e-3 * scale)          # parameters with a vanishing gradient: absolute floor
    return errs


def _report(tag, errs):
    v = np.array(sorted(errs.values()))
    worst = max(errs, key=errs.get)
    print("%s: gradient-norm relative error median %.2e  p95 %.2e  max %.2e (%s)"
          % (tag, np.median(v), v[int(0.95 * (len(v) - 1))], v[-1], worst))
    return v


def _fp64_yardstick(golden_dir, errs_vs64):
    """errs_vs64: per-parameter relative error of the product's gradient norms against the float64 reference run."""
    y = np.load(os.path.join(golden_dir, "detector_train_1024_first_fp64.npz"))
    ref = np.sort(y["ref_fp32_rel_dev"])
    v = np.array(sorted(errs_vs64.values()))
    q = lambda a, f: a[int(f * (len(a) - 1))]      # noqa: E731
    stats = [(np.median(v), np.median(ref)), (q(v, 0.95), q(ref, 0.95)), (v[-1], ref[-1])]
    print("   vs float64: product median %.2e p95 %.2e max %.2e | reference's own fp32: %.2e %.2e %.2e"
          % (stats[0][0], stats[1][0], stats[2][0], stats[0][1], stats[1][1], stats[2][1]))
    for got, own in stats:
        assert got <= 2.0 * own, (got, own)


def want_rows(key):
    return 4 if key.endswith("first4") else 8


def _grads_vs(golden_dir, named_grads, fname):
    g = np.load(os.path.join(golden_dir, fname))
    return _grad_norm_errors(named_grads, g)


@pytest.mark.parametrize("path", ["dynamic", "static"])
def test_fp32_step_matches_reference_at_baseline_size(golden_dir, monkeypatch, path):
    """fp32, deterministic sampler on both sides. dynamic = the reference-shaped path (model(list[dict])); static = the
    fixed-capacity path (4000 proposal slots, 512 ROI slots, 128 mask slots per image, validity masks)."""
    from u2seg_b200.modeling import rpn, static_train
    g, (n, H, W, K, S, seed, G), data = _golden(golden_dir, "first")
    model = _build_fp32(K, S)
    monkeypatch.setattr(rpn, "_randperm", _first)
    monkeypatch.setattr(static_train, "_rand_keys", _first_keys)
    if path == "dynamic":
        losses = model(_batch(data))
    else:
        packed = static_train.pack_batch(_batch(data), torch.device("cuda"), g_max=G)
        losses, flag = static_train.forward_train_static(model, *packed)
        assert not bool(flag)
    worst = _check_losses(losses, g, 1e-3)                                        # FP: within 1e-3 (fp32)
    print("%s fp32 2x%dx%d: worst loss error %.2e" % (path, H, W, worst))
    sum(losses.values()).backward()
    grads = {n_: p.grad for n_, p in model.named_parameters()}
    _report("%s fp32 vs the reference's fp32" % path, _grad_norm_errors(grads, g))
    _fp64_yardstick(golden_dir, _grads_vs(golden_dir, grads, "detector_train_1024_first_fp64.npz"))
    named = dict(model.named_parameters())
    y = np.load(os.path.join(golden_dir, "detector_train_1024_first_fp64.npz"))
    for key, name in (("grad_fpn_output3_first8", "backbone.fpn_output3.weight"),
                      ("grad_res4_0_conv1_first8", "backbone.bottom_up.res4.0.conv1.weight"),
                      ("grad_cls_score2_first4", "roi_heads.box_predictor.2.cls_score.weight")):
        # element-wise, against the float64 values, relative to the slice's largest entry; yardstick = the reference's own
        # float32 deviation on the same slice (stored), floor 1e-3
        got, want = named[name].grad[:want_rows(key)].double().cpu().numpy(), y[key]
        err = np.abs(got - want).max() / np.abs(want).max()
        own = float(y[key + "_ref_fp32_dev"])
        print("   %s[:%d] element-wise vs float64: product %.2e, reference's own fp32 %.2e" % (name, want_rows(key), err, own))
        assert err <= max(1e-3, 3.0 * own), (name, err, own)
    rm = model.state_dict()["backbone.bottom_up.stem.conv1.norm.running_mean"].cpu().numpy()
    np.testing.assert_allclose(rm, g["running_mean_stem"], rtol=1e-4, atol=1e-5)  # BN running statistics updated alike


def test_fp32_dynamic_step_with_reference_randperm_draws(golden_dir, monkeypatch):
    """The reference's own sampler: torch.randperm on the CPU generator, consumed in the reference's order. One borderline
    candidate changes n and with it the whole permutation, so only the losses are held to 1e-3 here (gradients: see the
    deterministic-sampler tests)."""
    from u2seg_b200.modeling import rpn
    g, (n, H, W, K, S, seed, G), data = _golden(golden_dir, "randperm")
    model = _build_fp32(K, S)
    monkeypatch.setattr(rpn, "_randperm", _cpu_randperm)
    torch.manual_seed(seed)
    losses = model(_batch(data))
    worst = _check_losses(losses, g, 1e-3)
    print("dynamic fp32 (randperm draws) 2x%dx%d: worst loss error %.2e" % (H, W, worst))
    sum(losses.values()).backward()
    v = _report("dynamic fp32 (randperm draws)", _grad_norm_errors({n_: p.grad for n_, p in model.named_parameters()}, g))
    assert v[int(0.5 * (len(v) - 1))] <= 2e-3 and v[-1] <= 1e-1


def test_bf16_static_graph_step_matches_reference_at_baseline_size(golden_dir, monkeypatch):
    """The benchmarked configuration: Trainer(static_graph=True, bf16 autocast) - tcgen05 convolutions, fused bf16
    SyncBN/GN, selected-class mask predictor, whole step replayed from one CUDA graph - on the BASELINE batch with the
    deterministic sampler. The RPN's top-k + NMS selection is discontinuous in the objectness logits (bf16 rounds them
    to 8 bits, reordering thousands of near-ties among 261,888 anchors per image), so the ROI heads are fed the
    reference's recorded proposals; everything else - backbone, FPN, semantic head, RPN head and its losses, the three
    cascade stages, the mask head, all gradients, the optimizer step at LR 0 - is the product's bf16 computation.
    Bounds: bf16 has an 8-bit mantissa (2^-9 = 2e-3 relative rounding per stored activation, ~100 layers deep): losses
    within 2e-2 of max(1,|v|); gradient norms within 5e-2 (median) / 2e-1 (95th percentile)."""
    from u2seg_b200.config import get_u2seg_cfg
    from u2seg_b200.engine import Trainer
    from u2seg_b200.modeling import rpn, static_train
    g, (n, H, W, K, S, seed, G), data = _golden(golden_dir, "first")
    monkeypatch.setattr(rpn, "_randperm", _first)
    monkeypatch.setattr(static_train, "_rand_keys", _first_keys)
    ref_props = torch.from_numpy(g["proposal_boxes"]).cuda()

    def recorded_proposals(rpn_mod, images_size, anchors, anchors_t, logits, deltas, flags):
        flags.append(torch.zeros((), dtype=torch.bool, device=ref_props.device))
        assert ref_props.shape[1] == rpn_mod.post_nms_topk[True]
        return ref_props, torch.ones(ref_props.shape[:2], dtype=torch.bool, device=ref_props.device)

    monkeypatch.setattr(static_train, "_rpn_proposals_static", recorded_proposals)
    cfg = get_u2seg_cfg(K)
    cfg.SOLVER.BASE_LR = 0.0
    cfg.SOLVER.WEIGHT_DECAY = 0.0
    cfg.SOLVER.WEIGHT_DECAY_NORM = 0.0
    tr = Trainer(cfg, amp_dtype=torch.bfloat16, static_graph=True, g_max=G)
    tr.load_state_dict(do.init_params(do.DetCfg(K, S), 0))
    losses = tr.run_step(_batch(data))
    torch.cuda.synchronize()
    tr.check_finite()
    worst = _check_losses(losses, g, 2e-2)
    print("static-graph bf16 2x%dx%d: worst loss error %.2e" % (H, W, worst))
    names = [n_ for n_, p in tr.model.named_parameters() if p.requires_grad]
    # the fused optimizer reads the flat fp32 gradient buffer and applies clipping on the fly: the buffer still holds
    # d(loss)/d(param)
    errs = _grad_norm_errors(dict(zip(names, tr._upd_grads)), g)
    v = _report("static-graph bf16", errs)
    assert v[int(0.5 * (len(v) - 1))] <= 5e-2 and v[int(0.95 * (len(v) - 1))] <= 2e-1, max(errs.items(), key=lambda kv: kv[1])
    # the step really ran on the tcgen05 kernels
    from u2seg_b200 import _lib
    assert tr.graph_own_launches > 300 and _lib.launch_count > 0
