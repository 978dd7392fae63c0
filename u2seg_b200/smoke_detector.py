"""__graft_entry__.smoke() leg for the detector path: one small PanopticFPN training step and one inference on
cuda:0 through the product model, checked against the CPU oracle (the oracle is the checker only)."""
import torch


def run():
    from oracle import detector_oracle as do

    from .config import get_u2seg_cfg
    from .modeling import build_model, rpn
    from .structures import BitMasks, Boxes, Instances

    K, S, seed = 800, 28, 5
    cfg_o = do.DetCfg(K, S)
    params = do.init_params(cfg_o, 0)
    data = do.synthetic_batch(1, 128, 160, K, S, seed=seed, G=4, min_size=16, max_size=80)
    torch.manual_seed(seed)
    want = do.forward_train(params, cfg_o, *data)

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model = build_model(get_u2seg_cfg(K))
    model.load_state_dict(params)
    model = model.to(memory_format=torch.channels_last).train()
    images, boxes, classes, masks, sems = data
    inst = Instances((128, 160))
    inst.gt_boxes, inst.gt_classes, inst.gt_masks = Boxes(boxes[0]), classes[0], BitMasks(masks[0])
    batch = [{"image": images[0], "instances": inst, "sem_seg": sems[0]}]
    saved = rpn._randperm
    rpn._randperm = lambda n, device=None: torch.randperm(n).to(device)   # the oracle's sampling order
    try:
        torch.manual_seed(seed)
        got = model(batch)
    finally:
        rpn._randperm = saved
    for k, v in want.items():
        g = float(got[k])
        assert abs(g - float(v)) <= 2e-3 * max(1.0, abs(float(v))), (k, g, float(v))
    sum(got.values()).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())
    model.eval()
    out = model([{"image": images[0], "height": 128, "width": 160}])[0]
    assert out["sem_seg"].shape == (S, 128, 160) and out["panoptic_seg"][0].shape == (128, 160)
    print("smoke: detector OK (10 training losses within 2e-3 of the oracle, backward finite, inference runs)")
