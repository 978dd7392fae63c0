"""Training / inference engine for the detector path — the B200-side counterpart of
detectron2/engine/{train_loop.py:479-521 AMPTrainer.run_step, defaults.py:60-79 create_ddp_model,
defaults.py:253-321 DefaultPredictor} and detectron2/solver/build.py:63-139 (SGD momentum 0.9 +
per-parameter gradient-norm clipping + WarmupMultiStepLR).

Data parallelism (one process per GPU): every parameter's .grad is a view into ONE flat fp32 buffer, so
the DDP-equivalent is a single NCCL all-reduce of that buffer per step (SUM / world), issued after
backward; SyncBN statistics are exchanged inside the model. Loss scalars stay on the device.
"""
import math

import torch
import torch.distributed as dist

from .modeling import build_model


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class WarmupMultiStepLR:
    """solver/build.py:283-323 + lr_scheduler.py (linear warmup, multi-step decay)."""

    def __init__(self, cfg):
        s = cfg.SOLVER
        self.base_lr, self.steps, self.gamma = s.BASE_LR, [x for x in s.STEPS if x <= s.MAX_ITER], s.GAMMA
        self.warmup_factor, self.warmup_iters = s.WARMUP_FACTOR, s.WARMUP_ITERS

    def lr(self, it):
        f = self.gamma ** sum(1 for s in self.steps if it >= s)
        if it < self.warmup_iters:
            alpha = it / self.warmup_iters
            f *= self.warmup_factor * (1 - alpha) + alpha
        return self.base_lr * f


class FlatGradients:
    """All gradients as views into one flat fp32 buffer -> the DDP-equivalent is ONE all-reduce per step."""

    def __init__(self, params, device):
        self.params = list(params)
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        if _world() > 1:
            dist.all_reduce(self.flat, group=group)
            self.flat.div_(_world())


class _BackboneTuple(torch.nn.Module):
    def __init__(self, backbone, keys):
        super().__init__()
        self.backbone, self.keys = backbone, keys

    def forward(self, x):
        out = self.backbone(x)
        return tuple(out[k] for k in self.keys)


class Trainer:
    """static_graph=True: the step runs through modeling/static_train.py (fixed-capacity device buffers, no host
    synchronisation) and forward + backward + clip + SGD are captured in ONE CUDA graph that is replayed every step;
    inputs are copied into static buffers. Requires same-size images and at most `g_max` instances per image."""

    def __init__(self, cfg, model=None, amp_dtype=torch.bfloat16, device=None, graph_backbone=False,
                 static_graph=False, g_max=None):
        self.cfg = cfg
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.model = model if model is not None else build_model(cfg)
        self.model = self.model.to(self.device).to(memory_format=torch.channels_last)
        self.model.train()
        self.amp_dtype = amp_dtype if cfg.SOLVER.AMP.ENABLED else None
        self.params = [p for p in self.model.parameters() if p.requires_grad]
        self.grads = FlatGradients(self.params, self.device)
        # solver/build.py:119-139 get_default_optimizer_params: norm layers get WEIGHT_DECAY_NORM
        s = cfg.SOLVER
        norm_ids = set()
        for m in self.model.modules():
            if isinstance(m, (torch.nn.modules.batchnorm._BatchNorm, torch.nn.GroupNorm)):
                norm_ids.update(id(p) for p in m.parameters(recurse=False))
        groups = [{"params": [p for p in self.params if id(p) in norm_ids], "weight_decay": s.WEIGHT_DECAY_NORM},
                  {"params": [p for p in self.params if id(p) not in norm_ids], "weight_decay": s.WEIGHT_DECAY}]
        self.optimizer = torch.optim.SGD(groups, lr=s.BASE_LR, momentum=s.MOMENTUM, nesterov=s.NESTEROV, foreach=True)
        self.clip = s.CLIP_GRADIENTS if s.CLIP_GRADIENTS.ENABLED else None
        self.sched = WarmupMultiStepLR(cfg)
        self.iter = 0
        self.scaler = torch.amp.GradScaler("cuda") if self.amp_dtype == torch.float16 else None
        self.graph_backbone = graph_backbone and _world() == 1
        self._graph_tried = False
        self.static_graph = static_graph
        self.g_max = g_max
        self._graph = None
        self._lr_t = torch.zeros((), dtype=torch.float32, device=self.device)
        self._mom_bufs = None
        self.nonfinite_flag = None

    # ---------------- static-shape, whole-step CUDA graph ----------------
    @torch.no_grad()
    def _sgd_foreach(self):
        """solver/build.py:119-139 SGD(momentum, weight decay) with the learning rate read from a device scalar, so
        that the captured graph follows the LR schedule."""
        s = self.cfg.SOLVER
        if self._mom_bufs is None:
            self._mom_bufs = [[torch.zeros_like(p) for p in g["params"]] for g in self.optimizer.param_groups]
        for g, bufs in zip(self.optimizer.param_groups, self._mom_bufs):
            params = g["params"]
            if not params:
                continue
            grads = [p.grad for p in params]
            if g["weight_decay"] != 0:
                torch._foreach_add_(grads, params, alpha=g["weight_decay"])
            torch._foreach_mul_(bufs, s.MOMENTUM)
            torch._foreach_add_(bufs, grads)
            upd = torch._foreach_mul(bufs, self._lr_t)
            torch._foreach_sub_(params, upd)

    @torch.no_grad()
    def _clip_foreach(self):
        if self.clip is None:
            return
        grads = [p.grad for p in self.params]
        norms = torch._foreach_norm(grads, self.clip.NORM_TYPE)
        coef = torch.clamp(self.clip.CLIP_VALUE / (torch.stack(norms) + 1e-6), max=1.0)
        torch._foreach_mul_(grads, list(coef.unbind(0)))

    def _static_step(self):
        from .modeling.static_train import forward_train_static
        self.grads.zero_()
        with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
            loss_dict, flag = forward_train_static(self.model, *self._static_in)
        sum(loss_dict.values()).backward()
        self.grads.all_reduce_mean()          # one NCCL all-reduce of the flat gradient buffer (captured with the graph)
        self._clip_foreach()
        self._sgd_foreach()
        return {k: v.detach() for k, v in loss_dict.items()}, flag

    def _load_static_inputs(self, batched_inputs):
        from .modeling.static_train import pack_batch
        g_max = self.g_max or max(1, max(len(d["instances"]) for d in batched_inputs))
        packed = pack_batch(batched_inputs, self.device, g_max)
        if getattr(self, "_static_in", None) is None:
            self.g_max = g_max
            self._static_in = [t.clone() for t in packed]
        else:
            for dst, src in zip(self._static_in, packed):
                assert dst.shape == src.shape, "static_graph needs constant input shapes (%s vs %s)" % (dst.shape, src.shape)
                dst.copy_(src, non_blocking=True)

    def _run_step_static(self, batched_inputs):
        self._lr_t.fill_(self.sched.lr(self.iter))
        self._load_static_inputs(batched_inputs)
        if self._graph is None:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            with torch.cuda.stream(side):     # warm-up outside capture: lazy initialisations, autotuning, workspaces
                for _ in range(3):
                    self._static_step()
            cur.wait_stream(side)
            torch.cuda.synchronize()
            from . import _lib
            self._graph = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count
            with torch.cuda.graph(self._graph):
                self._static_out = self._static_step()
            self.graph_own_launches = _lib.launch_count - l0      # libu2b200 kernels replayed per step
        self._graph.replay()
        from . import _lib
        _lib.count_launches(self.graph_own_launches)
        self.iter += 1
        losses, self.nonfinite_flag = self._static_out
        return losses

    def check_finite(self):
        """proposal_utils.py:105-110 divergence guard, read off the critical path (costs one host sync)."""
        if self.nonfinite_flag is not None and bool(self.nonfinite_flag):
            raise FloatingPointError("Predicted boxes or scores contain Inf/NaN. Training has diverged.")

    def _maybe_capture_backbone(self, batched_inputs):
        """CUDA-graph the forward+backward of the static-shape backbone (launch-bound inner loop of the step)."""
        if not self.graph_backbone or self._graph_tried:
            return
        self._graph_tried = True
        with torch.no_grad():
            x = self.model.preprocess_image(batched_inputs).tensor
        keys = list(self.model.backbone._out_features)
        wrapper = _BackboneTuple(self.model.backbone, keys)
        sample = torch.zeros_like(x)
        if self.amp_dtype is not None:
            with torch.autocast("cuda", dtype=self.amp_dtype, cache_enabled=False):
                graphed = torch.cuda.make_graphed_callables(wrapper, (sample,))
        else:
            graphed = torch.cuda.make_graphed_callables(wrapper, (sample,))
        self.model._graphed_backbone = (graphed, keys, tuple(x.shape))

    def run_step(self, batched_inputs):
        """train_loop.py:479-521. Returns the dict of (detached, device-resident) losses."""
        if self.static_graph:
            return self._run_step_static(batched_inputs)
        lr = self.sched.lr(self.iter)
        for g in self.optimizer.param_groups:
            g["lr"] = lr
        self._maybe_capture_backbone(batched_inputs)
        self.grads.zero_()
        if self.amp_dtype is not None:
            with torch.autocast("cuda", dtype=self.amp_dtype, cache_enabled=not self.graph_backbone):
                loss_dict = self.model(batched_inputs)
        else:
            loss_dict = self.model(batched_inputs)
        losses = sum(loss_dict.values())
        if self.scaler is not None:
            self.scaler.scale(losses).backward()
        else:
            losses.backward()
        self.grads.all_reduce_mean()   # the single gradient all-reduce (DDP: SUM / world)
        if self.scaler is not None:
            self.scaler.unscale_(self.optimizer)
        if self.clip is not None:   # solver/build.py:63-73: clip_grad_norm_ per parameter tensor
            assert self.clip.CLIP_TYPE == "norm"
            grads = [p.grad for p in self.params]
            norms = torch._foreach_norm(grads, self.clip.NORM_TYPE)
            coef = torch.clamp(self.clip.CLIP_VALUE / (torch.stack(norms) + 1e-6), max=1.0)
            torch._foreach_mul_(grads, list(coef.unbind(0)))
        if self.scaler is not None:
            self.scaler.step(self.optimizer)
            self.scaler.update()
        else:
            self.optimizer.step()
        self.iter += 1
        return {k: v.detach() for k, v in loss_dict.items()}


class DefaultPredictor:
    """engine/defaults.py:253-321 (single image in, post-processed dict out)."""

    def __init__(self, cfg, model=None):
        self.cfg = cfg
        self.model = (model if model is not None else build_model(cfg)).to(memory_format=torch.channels_last)
        self.model.eval()

    @torch.no_grad()
    def __call__(self, image_chw, height=None, width=None):
        d = {"image": image_chw}
        if height is not None:
            d["height"], d["width"] = height, width
        return self.model([d])[0]
