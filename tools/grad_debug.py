"""Developer tool: per-loss gradient comparison product (GPU) vs oracle (CPU)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from oracle import detector_oracle as do
from test_model_gpu import _build, _make_batch, _cpu_randperm
from u2seg_b200.modeling import rpn

K, S, seed = 800, 28, 5
cfg = do.DetCfg(K, S)
params = do.init_params(cfg, 0)
data = do.synthetic_batch(1, 128, 160, K, S, seed=seed, G=4, min_size=16, max_size=80)
names = ["backbone.fpn_output2.weight", "backbone.fpn_output5.weight", "backbone.bottom_up.res3.0.conv2.weight",
         "roi_heads.box_head.0.fc1.weight", "roi_heads.box_head.1.fc1.weight", "roi_heads.box_head.2.fc1.weight",
         "roi_heads.mask_head.mask_fcn2.weight", "proposal_generator.rpn_head.conv.weight",
         "sem_seg_head.p4.2.weight", "sem_seg_head.p2.0.weight", "backbone.bottom_up.stem.conv1.norm.weight"]
op = {k: v.clone().requires_grad_(k in names) for k, v in params.items()}
torch.manual_seed(seed)
ol = do.forward_train(op, cfg, *data)
model = _build(K, params, True)
rpn._randperm = _cpu_randperm
torch.manual_seed(seed)
ml = model(_make_batch(data))
named = dict(model.named_parameters())
for key in ol:
    print("==", key, float(ol[key]), float(ml[key]))
    og = torch.autograd.grad(ol[key], [op[n] for n in names], retain_graph=True, allow_unused=True)
    mg = torch.autograd.grad(ml[key], [named[n] for n in names], retain_graph=True, allow_unused=True)
    for n, a, b in zip(names, mg, og):
        if a is None and b is None:
            continue
        if a is None or b is None:
            print("   ", n, "None mismatch", a is None, b is None); continue
        a = a.float().cpu()
        d = float(b.abs().max()) + 1e-20
        print("    %-45s rel %.2e  (scale %.2e)" % (n, float((a - b).abs().max()) / d, d))
