"""Golden vectors for the detector path, produced by the UNMODIFIED reference (imported from
/root/reference through oracle/ref_stubs). Called from oracle/make_golden.py; build container only.
"""
import os

import numpy as np
import torch

from oracle import detector_oracle as do


def build_reference_model(num_classes=800, training=True):
    from detectron2.config import get_cfg
    from detectron2.modeling import build_model
    cfg = get_cfg()
    cfg.merge_from_file("/root/reference/configs/COCO-PanopticSegmentation/u2seg_R50_%d.yaml" % num_classes)
    cfg.MODEL.DEVICE = "cpu"
    model = build_model(cfg)
    model.train(training)
    return cfg, model


def reference_inputs(images, boxes, classes, masks, sems, train=True, out_sizes=None):
    from detectron2.structures import BitMasks, Boxes, Instances
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
        else:
            h, w = out_sizes[i] if out_sizes is not None else (im.shape[1], im.shape[2])
            d["height"], d["width"] = h, w
        batch.append(d)
    return batch


def gen_model(gold_dir):
    from detectron2.utils.events import EventStorage
    cfgo = do.DetCfg(num_classes=800)
    params = do.init_params(cfgo, seed=0)
    cfg, model = build_reference_model(800, True)
    sd = model.state_dict()
    missing = [k for k in sd if k not in params]
    extra = [k for k in params if k not in sd]
    assert not missing and not extra, (missing[:5], extra[:5])
    for k in sd:
        assert tuple(sd[k].shape) == tuple(params[k].shape), (k, sd[k].shape, params[k].shape)
    model.load_state_dict(params)

    # ---- training losses on 2 x 256x320 seeded synthetic images ----
    H, W, seed = 256, 320, 7
    data = do.synthetic_batch(2, H, W, 800, 28, seed=seed, G=6, min_size=24, max_size=160)
    batch = reference_inputs(*data, train=True)
    torch.manual_seed(seed)
    with EventStorage():
        losses = model(batch)
    ref_losses = {k: float(v) for k, v in losses.items()}
    print("reference losses", ref_losses)
    # gradients of the summed loss (what AMPTrainer.run_step backpropagates): per-parameter L2 norms + one full tensor
    sum(losses.values()).backward()
    gnames = sorted(n for n, p in model.named_parameters() if p.grad is not None)
    gnorms = np.array([float(dict(model.named_parameters())[n].grad.norm()) for n in gnames], dtype=np.float64)
    gfull = dict(model.named_parameters())["backbone.fpn_output3.weight"].grad[:8].numpy().copy()
    torch.manual_seed(seed)
    op = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in params.items()}
    ora = do.forward_train(op, cfgo, *data)
    print("oracle    losses", {k: float(v) for k, v in ora.items()})
    sum(ora.values()).backward()
    worst = max(abs(float(op[n].grad.norm()) - g) / (g + 1e-12) for n, g in zip(gnames, gnorms))
    print("oracle vs reference gradient norms: worst relative difference %.3e over %d parameters" % (worst, len(gnames)))
    np.savez_compressed(os.path.join(gold_dir, "detector_train_256x320.npz"),
                        keys=np.array(list(ref_losses.keys())), values=np.array(list(ref_losses.values()), dtype=np.float64),
                        grad_names=np.array(gnames), grad_norms=gnorms, grad_fpn_output3_first8=gfull,
                        meta=np.array([2, H, W, 800, 28, seed, 6, 24, 160], dtype=np.int64))

    # ---- inference on 1 x 200x304 image, output size 240x360 ----
    model.eval()
    H, W, seed = 200, 304, 11
    data = do.synthetic_batch(1, H, W, 800, 28, seed=seed, G=4, min_size=24, max_size=120)
    model.load_state_dict(do.eval_fixture_params(cfgo, data[0], seed=0))
    batch = reference_inputs(*data, train=False, out_sizes=[(240, 360)])
    with torch.no_grad():
        out = model(batch)[0]
    inst = out["instances"]
    pan, info = out["panoptic_seg"]
    np.savez_compressed(
        os.path.join(gold_dir, "detector_infer_200x304.npz"),
        pred_boxes=inst.pred_boxes.tensor.numpy(), scores=inst.scores.numpy(),
        pred_classes=inst.pred_classes.numpy(), pred_masks=np.packbits(inst.pred_masks.numpy(), axis=None),
        mask_shape=np.array(inst.pred_masks.shape), sem_seg_argmax=out["sem_seg"].argmax(0).numpy().astype(np.uint8),
        sem_seg_sample=out["sem_seg"][:, ::16, ::16].numpy(), panoptic=pan.numpy().astype(np.int32),
        n_segments=np.array([len(info)]), meta=np.array([1, H, W, 800, 28, seed, 4, 24, 120, 240, 360], dtype=np.int64))
    print("reference inference: %d detections, %d panoptic segments" % (len(inst), len(info)))


def gen_ops(gold_dir):
    """Op-level vectors from the reference's own layers/structures (and the golden values its tests hold)."""
    from detectron2.layers import ROIAlign, batched_nms, paste_masks_in_image
    from detectron2.modeling.anchor_generator import DefaultAnchorGenerator
    from detectron2.modeling.box_regression import Box2BoxTransform
    from detectron2.modeling.matcher import Matcher
    from detectron2.modeling.poolers import ROIPooler, assign_boxes_to_levels
    from detectron2.layers import ShapeSpec
    from detectron2.structures import BitMasks, Boxes, pairwise_iou
    g = torch.Generator().manual_seed(123)
    out = {}

    def rand_boxes(n, W, H, lo=4.0, hi=None):
        hi = hi or W / 2
        cx, cy = torch.rand(n, generator=g) * W, torch.rand(n, generator=g) * H
        w = torch.exp(torch.rand(n, generator=g) * (np.log(hi) - np.log(lo)) + np.log(lo))
        h = torch.exp(torch.rand(n, generator=g) * (np.log(hi) - np.log(lo)) + np.log(lo))
        b = torch.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1)
        b[:, 0::2] = b[:, 0::2].clamp(0, W)
        b[:, 1::2] = b[:, 1::2].clamp(0, H)
        return b

    # tests/layers/test_roi_align.py:14-47 — the reference's own golden case (5x5 arange, aligned / legacy)
    inp = torch.arange(25).reshape(1, 1, 5, 5).float()
    rois = torch.tensor([[0, 1, 1, 3, 3]], dtype=torch.float32)
    out["roialign_test_aligned"] = ROIAlign((4, 4), 1.0, 0, aligned=True)(inp, rois).numpy()
    # multi-level pooler on random features (poolers.py:206-263)
    feats = [torch.randn(2, 16, 64 // s, 96 // s, generator=g) for s in (1, 2, 4, 8)]   # strides 4,8,16,32 of a 256x384 image
    boxes = [rand_boxes(40, 384, 256, 4, 300), rand_boxes(25, 384, 256, 4, 300)]
    boxes[0][0] = torch.tensor([10.0, 10.0, 10.0 + 224.0, 10.0 + 224.0])     # exact canonical size
    boxes[0][1] = torch.tensor([0.0, 0.0, 112.0, 112.0])                     # exact power-of-two boundaries
    boxes[0][2] = torch.tensor([5.0, 5.0, 5.0, 9.0])                         # zero-area box
    pooler7 = ROIPooler(7, (0.25, 0.125, 0.0625, 0.03125), 0, "ROIAlignV2")
    pooler14 = ROIPooler(14, (0.25, 0.125, 0.0625, 0.03125), 0, "ROIAlignV2")
    out["pool_feat0"], out["pool_feat1"], out["pool_feat2"], out["pool_feat3"] = [f.numpy() for f in feats]
    out["pool_boxes0"], out["pool_boxes1"] = boxes[0].numpy(), boxes[1].numpy()
    out["pool_out7"] = pooler7(feats, [Boxes(b) for b in boxes]).numpy()
    out["pool_out14"] = pooler14(feats, [Boxes(b) for b in boxes]).numpy()
    out["pool_levels"] = assign_boxes_to_levels([Boxes(b) for b in boxes], 2, 5, 224, 4).numpy()
    # paste_masks_in_image (mask_ops.py:74-147)
    pm = torch.rand(9, 28, 28, generator=g)
    pb = rand_boxes(9, 150, 100, 6, 90)
    pb[0] = torch.tensor([-20.0, -10.0, 60.0, 50.0])          # partly outside the image
    out["paste_masks"], out["paste_boxes"] = pm.numpy(), pb.numpy()
    out["paste_out"] = np.packbits(paste_masks_in_image(pm, pb, (100, 150), 0.5).numpy(), axis=None)
    # pairwise_iou + Matcher (boxes.py:336, matcher.py:62) incl. the golden case of tests/modeling/test_matcher.py:16-24
    gt = rand_boxes(7, 384, 256, 20, 200)
    an = rand_boxes(3000, 384, 256, 8, 300)
    iou = pairwise_iou(Boxes(gt), Boxes(an))
    out["iou_gt"], out["iou_an"], out["iou"] = gt.numpy(), an.numpy(), iou.numpy()
    m, l = Matcher([0.3, 0.7], [0, -1, 1], allow_low_quality_matches=True)(iou)
    out["match_rpn_idx"], out["match_rpn_lab"] = m.numpy(), l.numpy()
    m, l = Matcher([0.5], [0, 1], allow_low_quality_matches=False)(iou)
    out["match_roi_idx"], out["match_roi_lab"] = m.numpy(), l.numpy()
    # anchors (anchor_generator.py:218) for a 64x96 image pyramid
    ag = DefaultAnchorGenerator(sizes=[[32], [64], [128], [256], [512]], aspect_ratios=[[0.5, 1.0, 2.0]],
                                strides=[4, 8, 16, 32, 64], offset=0.0)
    anc = ag([torch.zeros(1, 1, 64 // s, 96 // s) for s in (1, 2, 4, 8, 16)])
    for i, a in enumerate(anc):
        out["anchors%d" % i] = a.tensor.numpy()
    # box transform (box_regression.py:43-116)
    tr = Box2BoxTransform(weights=(10.0, 10.0, 5.0, 5.0))
    src, dst = rand_boxes(50, 384, 256, 8, 200), rand_boxes(50, 384, 256, 8, 200)
    d = tr.get_deltas(src, dst)
    out["b2b_src"], out["b2b_dst"], out["b2b_deltas"] = src.numpy(), dst.numpy(), d.numpy()
    big = d.clone()
    big[:5, 2:] += 9.0          # exercise the scale clamp
    out["b2b_big"], out["b2b_applied"] = big.numpy(), tr.apply_deltas(big, src).numpy()
    # batched_nms (layers/nms.py:9) — >1000 boxes so torchvision CPU takes its class-by-class path
    nb = rand_boxes(3000, 384, 256, 8, 120)
    ns = torch.rand(3000, generator=g)
    ni = torch.randint(0, 5, (3000,), generator=g)
    out["nms_boxes"], out["nms_scores"], out["nms_idxs"] = nb.numpy(), ns.numpy(), ni.numpy()
    out["nms_keep_065"] = batched_nms(nb, ns, ni, 0.65).numpy()
    out["nms_keep_050"] = batched_nms(nb, ns, ni, 0.5).numpy()
    # BitMasks.crop_and_resize (masks.py:191-222)
    ys, xs = torch.arange(100)[:, None] + 0.5, torch.arange(150)[None, :] + 0.5
    gm = torch.stack([((xs - 40 - 10 * i) / (20 + 3 * i)) ** 2 + ((ys - 50) / (15 + 2 * i)) ** 2 <= 1 for i in range(6)])
    cb = rand_boxes(6, 150, 100, 10, 120)
    out["crop_masks"] = np.packbits(gm.numpy(), axis=None)
    out["crop_boxes"] = cb.numpy()
    out["crop_out"] = np.packbits(BitMasks(gm).crop_and_resize(cb, 28).numpy(), axis=None)
    np.savez_compressed(os.path.join(gold_dir, "detector_ops.npz"), **out)
    print("wrote detector_ops.npz with", len(out), "arrays")


def gen_model_baseline(gold_dir):
    """BASELINE.json config 2: u2seg_R50_800 training step on 2 x 1024x1024 synthetic images with G=20 instances
    (SURVEY 8(d) recipe), run through the UNMODIFIED reference on CPU (fp32). Two fixtures:
      * sampler "randperm": the reference's own torch.randperm draws after torch.manual_seed(seed) (sampling.py:44-47);
        the product's dynamic path consumes the same CPU permutations (tests inject rpn._randperm);
      * sampler "first": torch.randperm replaced by arange for the duration of the forward, i.e. "first k candidates
        in index order" - the only deterministic rule the fixed-capacity static path can reproduce exactly.
    Stored: the 10 losses, per-parameter gradient L2 norms of all 248 parameters, and two small gradient slices.
    """
    import time
    from detectron2.utils.events import EventStorage
    cfgo = do.DetCfg(num_classes=800)
    params = do.init_params(cfgo, seed=0)
    cfg, model = build_reference_model(800, True)
    H, W, seed, G, lo, hi = 1024, 1024, 1234, 20, 32, 512
    data = do.synthetic_batch(2, H, W, 800, 28, seed=seed, G=G, min_size=lo, max_size=hi)
    batch = reference_inputs(*data, train=True)
    real_randperm = torch.randperm
    captured = {}
    model.proposal_generator.register_forward_hook(lambda mod, inp, out: captured.__setitem__("proposals", out[0]))
    for sampler in ("randperm", "first"):
        model.load_state_dict(params)       # also resets the BN running statistics
        model.zero_grad(set_to_none=True)
        if sampler == "first":
            torch.randperm = lambda n, *a, device=None, **k: torch.arange(n, device=device)
        try:
            torch.manual_seed(seed)
            t0 = time.time()
            with EventStorage():
                losses = model(batch)
            sum(losses.values()).backward()
        finally:
            torch.randperm = real_randperm
        ref_losses = {k: float(v) for k, v in losses.items()}
        print("reference [%s] %.1f s" % (sampler, time.time() - t0), ref_losses)
        named = dict(model.named_parameters())
        gnames = sorted(n for n, p in named.items() if p.grad is not None)
        gnorms = np.array([float(named[n].grad.double().norm()) for n in gnames], dtype=np.float64)
        np.savez_compressed(
            os.path.join(gold_dir, "detector_train_1024_%s.npz" % sampler),
            keys=np.array(list(ref_losses.keys())), values=np.array(list(ref_losses.values()), dtype=np.float64),
            grad_names=np.array(gnames), grad_norms=gnorms,
            grad_fpn_output3_first8=named["backbone.fpn_output3.weight"].grad[:8].numpy().copy(),
            grad_res4_0_conv1_first8=named["backbone.bottom_up.res4.0.conv1.weight"].grad[:8].numpy().copy(),
            grad_cls_score2_first4=named["roi_heads.box_predictor.2.cls_score.weight"].grad[:4].numpy().copy(),
            running_mean_stem=model.state_dict()["backbone.bottom_up.stem.conv1.norm.running_mean"].numpy().copy(),
            # the RPN's output (rpn.py:431-480): post-NMS proposals per image in score order, before GT boxes are appended.
            # Lets a reduced-precision run be compared downstream of the (discontinuous) top-k / NMS selection.
            proposal_boxes=np.stack([pr.proposal_boxes.tensor.numpy() for pr in captured["proposals"]]),
            meta=np.array([2, H, W, 800, 28, seed, G, lo, hi], dtype=np.int64))


def gen_model_baseline_fp64(gold_dir):
    """Rounding-noise yardstick for the BASELINE-config fixture: the same unmodified reference step ('first' sampler) run
    in float64 on CPU. Stored next to the fp32 values: losses and per-parameter gradient norms in fp64, and the reference's
    OWN fp32-vs-fp64 deviation per parameter - the precision any fp32 implementation of this step can be held to (the
    backward pass goes through 61 batch-norm layers; some gradient norms are reproducible only to a few 1e-3 in fp32)."""
    import time
    from detectron2.utils.events import EventStorage
    cfgo = do.DetCfg(num_classes=800)
    params = do.init_params(cfgo, seed=0)
    cfg, model = build_reference_model(800, True)
    H, W, seed, G, lo, hi = 1024, 1024, 1234, 20, 32, 512
    data = do.synthetic_batch(2, H, W, 800, 28, seed=seed, G=G, min_size=lo, max_size=hi)
    batch = reference_inputs(*data, train=True)
    for d in batch:
        d["instances"].gt_boxes.tensor = d["instances"].gt_boxes.tensor.double()
    model.load_state_dict(params)
    model = model.double()
    real_randperm = torch.randperm
    torch.randperm = lambda n, *a, device=None, **k: torch.arange(n, device=device)
    # layers/nms.py:20 calls torchvision with boxes.float() and the scores as they are: torchvision wants equal dtypes,
    # so the float64 run hands it float32 scores too (selection only; nothing differentiable passes through NMS)
    import importlib
    d2nms = importlib.import_module("detectron2.layers.nms")
    real_bnms = d2nms.box_ops.batched_nms
    d2nms.box_ops.batched_nms = lambda b, sc, i, t: real_bnms(b.float(), sc.float(), i, t)
    try:
        torch.manual_seed(seed)
        t0 = time.time()
        with EventStorage():
            losses = model(batch)
        sum(losses.values()).backward()
    finally:
        torch.randperm = real_randperm
        d2nms.box_ops.batched_nms = real_bnms
    print("reference fp64 [first] %.1f s" % (time.time() - t0), {k: float(v) for k, v in losses.items()})
    g32 = np.load(os.path.join(gold_dir, "detector_train_1024_first.npz"))
    named = dict(model.named_parameters())
    gnames = [str(n) for n in g32["grad_names"]]
    gn64 = np.array([float(named[n].grad.norm()) for n in gnames], dtype=np.float64)
    dev = np.abs(g32["grad_norms"] - gn64) / np.maximum(gn64, 1e-300)
    l64 = np.array([float(losses[str(k)]) for k in g32["keys"]], dtype=np.float64)
    print("reference fp32 vs fp64: losses max rel %.2e; gradient norms median %.2e p95 %.2e max %.2e (%s)"
          % (np.max(np.abs(g32["values"] - l64) / np.maximum(1, np.abs(l64))), np.median(dev), np.sort(dev)[int(0.95 * (len(dev) - 1))],
             dev.max(), gnames[int(dev.argmax())]))
    slices = {"grad_fpn_output3_first8": named["backbone.fpn_output3.weight"].grad[:8].numpy().copy(),
              "grad_res4_0_conv1_first8": named["backbone.bottom_up.res4.0.conv1.weight"].grad[:8].numpy().copy(),
              "grad_cls_score2_first4": named["roi_heads.box_predictor.2.cls_score.weight"].grad[:4].numpy().copy()}
    # element-wise deviation of the reference's own fp32 gradients from the fp64 ones, relative to the slice's max |g|
    slice_dev = {k + "_ref_fp32_dev": float(np.abs(g32[k].astype(np.float64) - v).max() / np.abs(v).max()) for k, v in slices.items()}
    print("reference fp32 vs fp64, element-wise on the stored slices:", slice_dev)
    np.savez_compressed(os.path.join(gold_dir, "detector_train_1024_first_fp64.npz"), keys=g32["keys"], values=l64,
                        grad_names=g32["grad_names"], grad_norms=gn64, ref_fp32_rel_dev=dev, meta=g32["meta"],
                        **slices, **{k: np.array(v) for k, v in slice_dev.items()})
