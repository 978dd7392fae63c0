"""CPU: the detector oracle (oracle/detector_oracle.py) is pinned against outputs of the unmodified
reference model and layers recorded by oracle/make_golden_detector.py."""
import os

import numpy as np
import pytest
import torch

from oracle import detector_oracle as do


@pytest.fixture(scope="module")
def ops(golden_dir):
    return np.load(os.path.join(golden_dir, "detector_ops.npz"))


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_roialign_reference_golden_case(ops):
    # the reference's own known-answer test: tests/layers/test_roi_align.py:14-47
    want = np.array([[4.5, 5.0, 5.5, 6.0], [7.0, 7.5, 8.0, 8.5], [9.5, 10.0, 10.5, 11.0], [12.0, 12.5, 13.0, 13.5]])
    assert np.allclose(ops["roialign_test_aligned"][0, 0], want)
    inp = torch.arange(25).reshape(1, 1, 5, 5).float()
    got = do.tv_roi_align(inp, torch.tensor([[0, 1, 1, 3, 3.0]]), (4, 4), 1.0, 0, True)
    assert np.allclose(got[0, 0].numpy(), want)


def test_pooler_and_levels(ops):
    feats = [T(ops["pool_feat%d" % i]) for i in range(4)]
    boxes = [T(ops["pool_boxes0"]), T(ops["pool_boxes1"])]
    assert np.array_equal(do.assign_levels(torch.cat(boxes)).numpy(), ops["pool_levels"])      # INT exact
    assert np.array_equal(do.roi_pool(feats, boxes, 7).numpy(), ops["pool_out7"])
    assert np.array_equal(do.roi_pool(feats, boxes, 14).numpy(), ops["pool_out14"])


def test_paste_masks(ops):
    got = do.paste_masks_in_image(T(ops["paste_masks"]), T(ops["paste_boxes"]), (100, 150), 0.5)
    assert np.array_equal(np.packbits(got.numpy(), axis=None), ops["paste_out"])


def test_iou_and_matcher(ops):
    iou = do.pairwise_iou(T(ops["iou_gt"]), T(ops["iou_an"]))
    assert np.array_equal(iou.numpy(), ops["iou"])
    m, l = do.matcher(iou, (0.3, 0.7), (0, -1, 1), True)
    assert np.array_equal(m.numpy(), ops["match_rpn_idx"]) and np.array_equal(l.numpy(), ops["match_rpn_lab"])
    m, l = do.matcher(iou, (0.5,), (0, 1), False)
    assert np.array_equal(m.numpy(), ops["match_roi_idx"]) and np.array_equal(l.numpy(), ops["match_roi_lab"])
    # the reference's own golden: tests/modeling/test_matcher.py:16-24
    q = torch.tensor([[0.15, 0.45, 0.2, 0.6], [0.3, 0.65, 0.05, 0.1], [0.05, 0.4, 0.25, 0.4]])
    m, l = do.matcher(q, (0.3, 0.5), (0, -1, 1), True)
    assert m.tolist() == [1, 1, 2, 0] and l.tolist() == [-1, 1, 0, 1]


def test_anchors(ops):
    anc = do.make_anchors([(64 // s, 96 // s) for s in (1, 2, 4, 8, 16)], do.DetCfg())
    for i, a in enumerate(anc):
        assert np.array_equal(a.numpy(), ops["anchors%d" % i])


def test_box_transform(ops):
    d = do.get_deltas(T(ops["b2b_src"]), T(ops["b2b_dst"]), (10.0, 10.0, 5.0, 5.0))
    assert np.array_equal(d.numpy(), ops["b2b_deltas"])
    a = do.apply_deltas(T(ops["b2b_big"]), T(ops["b2b_src"]), (10.0, 10.0, 5.0, 5.0), do.DetCfg().scale_clamp)
    assert np.array_equal(a.numpy(), ops["b2b_applied"])


def test_batched_nms(ops):
    b, s, i = T(ops["nms_boxes"]), T(ops["nms_scores"]), T(ops["nms_idxs"])
    assert np.array_equal(do.batched_nms(b, s, i, 0.65).numpy(), ops["nms_keep_065"])
    assert np.array_equal(do.batched_nms(b, s, i, 0.5).numpy(), ops["nms_keep_050"])


def test_crop_and_resize(ops):
    gm = T(np.unpackbits(ops["crop_masks"])[:6 * 100 * 150].reshape(6, 100, 150).astype(bool))
    got = do.crop_and_resize_masks(gm, T(ops["crop_boxes"]), 28)
    assert np.array_equal(np.packbits(got.numpy(), axis=None), ops["crop_out"])


def test_training_losses_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "detector_train_256x320.npz"))
    n, H, W, K, S, seed, G, lo, hi = [int(v) for v in g["meta"]]
    cfg = do.DetCfg(num_classes=K, sem_classes=S)
    params = do.init_params(cfg, seed=0)
    data = do.synthetic_batch(n, H, W, K, S, seed=seed, G=G, min_size=lo, max_size=hi)
    torch.manual_seed(seed)
    losses = do.forward_train(params, cfg, *data)
    assert list(losses.keys()) == [str(k) for k in g["keys"]]
    for k, v in zip(g["keys"], g["values"]):
        assert abs(float(losses[str(k)]) - v) <= 1e-6 * max(1.0, abs(v)), (k, float(losses[str(k)]), v)


def test_training_gradients_match_reference(golden_dir):
    """backward parity of the oracle: per-parameter gradient norms of the summed loss vs the reference's
    (catches forward-neutral pieces such as cascade_rcnn._ScaleGradient)."""
    g = np.load(os.path.join(golden_dir, "detector_train_256x320.npz"))
    n, H, W, K, S, seed, G, lo, hi = [int(v) for v in g["meta"]]
    cfg = do.DetCfg(num_classes=K, sem_classes=S)
    params = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k)
              for k, v in do.init_params(cfg, seed=0).items()}
    data = do.synthetic_batch(n, H, W, K, S, seed=seed, G=G, min_size=lo, max_size=hi)
    torch.manual_seed(seed)
    sum(do.forward_train(params, cfg, *data).values()).backward()
    assert len(g["grad_names"]) == 248
    for name, want in zip(g["grad_names"], g["grad_norms"]):
        got = float(params[str(name)].grad.norm())
        assert abs(got - want) <= 1e-4 * max(want, 1e-6), (str(name), got, want)
    np.testing.assert_allclose(params["backbone.fpn_output3.weight"].grad[:8].numpy(), g["grad_fpn_output3_first8"],
                               rtol=1e-4, atol=1e-7)


def test_inference_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "detector_infer_200x304.npz"))
    n, H, W, K, S, seed, G, lo, hi, oh, ow = [int(v) for v in g["meta"]]
    cfg = do.DetCfg(num_classes=K, sem_classes=S)
    data = do.synthetic_batch(n, H, W, K, S, seed=seed, G=G, min_size=lo, max_size=hi)
    params = do.eval_fixture_params(cfg, data[0], seed=0)
    out = do.forward_inference(params, cfg, data[0], out_sizes=[(oh, ow)])[0]
    inst = out["instances"]
    assert np.array_equal(inst["pred_classes"].numpy(), g["pred_classes"])        # INT exact
    np.testing.assert_allclose(inst["pred_boxes"].numpy(), g["pred_boxes"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(inst["scores"].numpy(), g["scores"], rtol=1e-5, atol=1e-7)
    assert np.array_equal(np.packbits(inst["pred_masks"].numpy(), axis=None), g["pred_masks"])
    assert np.array_equal(out["sem_seg"].argmax(0).numpy().astype(np.uint8), g["sem_seg_argmax"])
    assert np.array_equal(out["panoptic_seg"][0].numpy(), g["panoptic"])
    assert len(out["panoptic_seg"][1]) == int(g["n_segments"][0])


@pytest.mark.reference
def test_oracle_param_names_match_reference_state_dict():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "oracle", "ref_stubs"), "/root/reference"]
    from oracle.make_golden_detector import build_reference_model
    for K in (300, 800):
        _, model = build_reference_model(K, True)
        sd = model.state_dict()
        p = do.init_params(do.DetCfg(num_classes=K), seed=0)
        assert set(sd.keys()) == set(p.keys())
        assert all(tuple(sd[k].shape) == tuple(p[k].shape) for k in sd)
