"""Training / inference engine for the detector path — the B200-side counterpart of
detectron2/engine/{train_loop.py:479-521 AMPTrainer.run_step, defaults.py:60-79 create_ddp_model,
defaults.py:253-321 DefaultPredictor} and detectron2/solver/build.py:63-139 (SGD momentum 0.9 +
per-parameter gradient-norm clipping + WarmupMultiStepLR).

Data parallelism (one process per GPU): every parameter's .grad is a view into ONE flat fp32 buffer, so
the DDP-equivalent is a single NCCL all-reduce of that buffer per step (SUM / world), issued after
backward; SyncBN statistics are exchanged inside the model. Loss scalars stay on the device.
"""
import math

import os

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


def _like_view(flat, off, p):
    """A view of flat[off : off+p.numel()] with p's shape AND strides (p dense: contiguous or channels_last)."""
    return torch.as_strided(flat, p.shape, p.stride(), off)


def _aligned(n, a=64):
    """Every tensor starts on a 64-element boundary of its flat buffer (>= 128 B: TMA descriptors of the tcgen05 conv
    need 16-byte aligned bases, vectorised multi-tensor kernels 16 B too); the padding stays zero."""
    return (n + a - 1) // a * a


class FlatGradients:
    """All gradients as views into one flat fp32 buffer -> the DDP-equivalent is ONE all-reduce per step.
    `tail` lists tensors whose fp32 gradients live at the end of the buffer without being attached as `.grad`
    (the bf16 compute weights of the static-graph trainer: their bf16 `.grad` is converted into that tail)."""

    def __init__(self, params, device, tail=()):
        self.params = list(params)
        self.n_head = sum(_aligned(p.numel()) for p in self.params)
        total = self.n_head + sum(_aligned(p.numel()) for p in tail)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        off = 0
        for p in self.params:
            p.grad = _like_view(self.flat, off, p)
            off += _aligned(p.numel())
        self.tail_views = []
        for p in tail:
            self.tail_views.append(_like_view(self.flat, off, p))
            off += _aligned(p.numel())

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
        self.lowp = static_graph and self.amp_dtype == torch.bfloat16 and os.environ.get("U2B_BF16_WEIGHTS", "1") != "0"
        if self.lowp:
            self._setup_lowp_weights()
        else:
            self.grads = FlatGradients(self.params, self.device)
            self._upd_params, self._upd_grads = self.params, None
            self._flat_views = [p.grad for p in self.params]
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
    def _setup_lowp_weights(self):
        """Conv / linear parameters are consumed in bf16 under autocast (amp: train_loop.py:479-521). Instead of one
        fp32->bf16 cast kernel per parameter per step (and one bf16->fp32 cast per weight gradient), the modules hold
        bf16 COMPUTE copies (views of one flat bf16 buffer, refreshed by ONE kernel after the optimizer step) and the
        fp32 MASTER weights live in a flat buffer owned by the trainer; the bf16 weight gradients are widened into the
        tail of the flat fp32 gradient buffer by the same multi-tensor copy that gathers all gradients.
        Numerically identical to autocast: same rounding of the same fp32 masters, same bf16 wgrad outputs."""
        low_ids = set()
        for m in self.model.modules():
            if isinstance(m, (torch.nn.modules.conv._ConvNd, torch.nn.Linear)):
                low_ids.update(id(p) for p in m.parameters(recurse=False) if p.requires_grad)
        self._low_params = [p for p in self.params if id(p) in low_ids]
        others = [p for p in self.params if id(p) not in low_ids]
        n = sum(_aligned(p.numel()) for p in self._low_params)
        self._master_flat = torch.zeros(n, dtype=torch.float32, device=self.device)
        self._w16_flat = torch.zeros(n, dtype=torch.bfloat16, device=self.device)
        self.grads = FlatGradients(others, self.device, tail=self._low_params)
        self._masters = {}
        off = 0
        for p in self._low_params:
            master = _like_view(self._master_flat, off, p)
            master.copy_(p.data)
            w16 = _like_view(self._w16_flat, off, p)
            w16.copy_(p.data)
            p.data = w16
            p.grad = None
            self._masters[id(p)] = master
            off += _aligned(p.numel())
        tail = {id(p): v for p, v in zip(self._low_params, self.grads.tail_views)}
        self._upd_params = [self._masters.get(id(p), p) for p in self.params]       # what SGD updates (fp32)
        self._upd_grads = [tail.get(id(p), p.grad) for p in self.params]            # fp32 gradients, same order

    def master_parameters(self):
        """fp32 parameters by name (the bf16 compute copies are an implementation detail of the static step)."""
        out = {}
        for (name, p), m in zip(((n, q) for n, q in self.model.named_parameters() if q.requires_grad), self._upd_params):
            out[name] = m
        return out

    @torch.no_grad()
    def broadcast_parameters(self, src=0):
        """DDP's initial broadcast: rank `src`'s parameters and buffers everywhere."""
        if _world() == 1:
            return
        for t in list(self._upd_params) + list(self.model.buffers()):
            dist.broadcast(t, src)
        if self.lowp:
            self._w16_flat.copy_(self._master_flat)

    def _groups(self):
        """(fp32 params, fp32 grads, weight decay) per optimizer group."""
        idx = {id(p): i for i, p in enumerate(self.params)}
        for g in self.optimizer.param_groups:
            ii = [idx[id(p)] for p in g["params"]]
            yield ([self._upd_params[i] for i in ii],
                   [self._upd_grads[i] if self._upd_grads is not None else self._flat_views[i] for i in ii],
                   g["weight_decay"])

    @torch.no_grad()
    def _sgd_foreach(self):
        """solver/build.py:119-139 SGD(momentum, weight decay) with the learning rate read from a device scalar, so
        that the captured graph follows the LR schedule."""
        s = self.cfg.SOLVER
        groups = list(self._groups())
        if self._mom_bufs is None:
            self._mom_bufs = [[torch.zeros_like(p) for p in params] for params, _, _ in groups]
        for (params, grads, wd), bufs in zip(groups, self._mom_bufs):
            if not params:
                continue
            if wd != 0:
                torch._foreach_add_(grads, params, alpha=wd)
            torch._foreach_mul_(bufs, s.MOMENTUM)
            torch._foreach_add_(bufs, grads)
            upd = torch._foreach_mul(bufs, self._lr_t)
            torch._foreach_sub_(params, upd)
        if self.lowp:
            self._w16_flat.copy_(self._master_flat)      # refresh the bf16 compute weights: one kernel

    @torch.no_grad()
    def _clip_foreach(self):
        if self.clip is None:
            return
        grads = self._upd_grads if self._upd_grads is not None else self._flat_views
        norms = torch._foreach_norm(grads, self.clip.NORM_TYPE)
        coef = torch.clamp(self.clip.CLIP_VALUE / (torch.stack(norms) + 1e-6), max=1.0)
        torch._foreach_mul_(grads, list(coef.unbind(0)))

    def _static_step(self):
        from .modeling.static_train import forward_train_static
        # .grad = None: autograd hands over each gradient tensor instead of launching one `grad += g` kernel per
        # parameter; one multi-tensor copy then gathers (and, for bf16 weight gradients, widens) them into the flat
        # fp32 buffer. Under graph capture the gradient tensors live in the graph's private pool.
        for p in self.params:
            p.grad = None
        with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
            loss_dict, flag = forward_train_static(self.model, *self._static_in)
        sum(loss_dict.values()).backward()
        dst = self._upd_grads if self._upd_grads is not None else self._flat_views
        have = [(d, p.grad) for d, p in zip(dst, self.params) if p.grad is not None]
        if len(have) != len(self.params):
            self.grads.zero_()                # parameters outside the graph of this step keep a zero gradient
        with torch.no_grad():
            # the multi-tensor fast path is all-or-nothing per call: keep same-dtype / same-layout pairs together
            buckets = {}
            for d, g in have:
                buckets.setdefault((g.dtype, d.stride() == g.stride()), ([], []))
                buckets[(g.dtype, d.stride() == g.stride())][0].append(d)
                buckets[(g.dtype, d.stride() == g.stride())][1].append(g)
            for dsts, srcs in buckets.values():
                torch._foreach_copy_(dsts, srcs)
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
