"""Training / inference engine for the detector path — the B200-side counterpart of
detectron2/engine/{train_loop.py:479-521 AMPTrainer.run_step, defaults.py:60-79 create_ddp_model,
defaults.py:253-321 DefaultPredictor} and detectron2/solver/build.py:63-139 (SGD momentum 0.9 +
per-parameter gradient-norm clipping + WarmupMultiStepLR).

Data parallelism (one process per GPU): all gradients live in ONE flat fp32 buffer (dynamic step: every .grad is a
view of it; static-graph step: autograd's gradients are gathered into it by a multi-tensor copy), so the
DDP-equivalent is a single NCCL all-reduce of that buffer per step (SUM / world), issued after backward; SyncBN
statistics are exchanged inside the model. Loss scalars stay on the device.

Two step implementations share the kernels: the reference-shaped eager step (`static_graph=False`, any shapes) and the
static-shape step replayed from one CUDA graph (`static_graph=True`: flat fp32 masters + bf16 compute weights, fused
clip + SGD kernel, optional prefetch of the next batch on a copy stream).
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
    return torch.as_strided(flat, p.shape, p.stride(), flat.storage_offset() + off)   # as_strided offsets are absolute


def _aligned(n, a=64):
    """Every tensor starts on a 64-element boundary of its flat buffer (>= 128 B: TMA descriptors of the tcgen05 conv
    need 16-byte aligned bases, vectorised multi-tensor kernels 16 B too); the padding stays zero."""
    return (n + a - 1) // a * a


def _all_reduce_avg(t, group=None):
    """DDP's gradient averaging: SUM / world in one NCCL call (ReduceOp.AVG) - no separate division pass."""
    try:
        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=group)
    except (RuntimeError, ValueError):      # backends without AVG (gloo)
        dist.all_reduce(t, group=group)
        t.div_(_world())


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
            _all_reduce_avg(self.flat, group)


class _BackboneTuple(torch.nn.Module):
    def __init__(self, backbone, keys):
        super().__init__()
        self.backbone, self.keys = backbone, keys

    def forward(self, x):
        out = self.backbone(x)
        return tuple(out[k] for k in self.keys)


class Trainer:
    """train_loop.py:479-521 run_step + solver/build.py optimizer, on the B200 kernels.

    static_graph=True: the step runs through modeling/static_train.py (fixed-capacity device buffers, no host
    synchronisation) and forward + backward + all-reduce + clip + SGD are captured in ONE CUDA graph that is replayed
    every step; inputs are copied into static buffers (`prefetch` stages the next batch meanwhile). Requires same-size
    images and at most `g_max` instances per image. With bf16 autocast the conv / linear parameters become bf16
    compute copies of fp32 masters held by the trainer (`master_parameters()` returns the fp32 values by name).
    graph_backbone=True (eager step only): CUDA-graph just the static-shape backbone + FPN."""

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
        if static_graph:      # fixed input shapes on every rank: the in-kernel SyncBN exchange may assume equal counts
            from .modeling import fused_bn
            fused_bn.EQUAL_SHAPES_ACROSS_RANKS = True
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
        self.grads = FlatGradients(others, self.device, tail=self._low_params)
        n_head = self.grads.n_head
        # ONE fp32 master buffer for every parameter, laid out exactly like the flat gradient buffer: first the
        # parameters that stay fp32 (norm layers; the modules' tensors become views of it), then the masters of the
        # bf16 compute weights. Momentum uses the same offsets (csrc/optimizer.cu walks the three in lockstep).
        self._master_all = torch.zeros(n_head + n, dtype=torch.float32, device=self.device)
        self._mom_all = torch.zeros_like(self._master_all)
        self._master_flat = self._master_all[n_head:]
        self._w16_flat = torch.zeros(n, dtype=torch.bfloat16, device=self.device)
        off = 0
        for p in others:
            v = _like_view(self._master_all, off, p)
            v.copy_(p.data)
            p.data = v
            off += _aligned(p.numel())
        assert off == n_head
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
        # segment tables of the fused optimizer kernel: parameters in buffer order
        seg_params = others + self._low_params
        norm_ids = set()
        for m in self.model.modules():
            if isinstance(m, (torch.nn.modules.batchnorm._BatchNorm, torch.nn.GroupNorm)):
                norm_ids.update(id(p) for p in m.parameters(recurse=False))
        sv = self.cfg.SOLVER
        chunk_seg = torch.full((self._master_all.numel() // 64,), -1, dtype=torch.int32)
        off = 0
        for i, p in enumerate(seg_params):
            chunk_seg[off // 64:(off + _aligned(p.numel())) // 64] = i
            off += _aligned(p.numel())
        self._seg_chunk = chunk_seg.to(self.device)
        self._seg_wd = torch.tensor([sv.WEIGHT_DECAY_NORM if id(p) in norm_ids else sv.WEIGHT_DECAY for p in seg_params],
                                    dtype=torch.float32, device=self.device)
        by_id = {id(p): g for p, g in zip(self.params, self._upd_grads)}
        self._seg_grads = [by_id[id(p)] for p in seg_params]                         # fp32 gradient views, buffer order

    def master_parameters(self):
        """fp32 parameters by name (the bf16 compute copies are an implementation detail of the static step)."""
        out = {}
        for (name, p), m in zip(((n, q) for n, q in self.model.named_parameters() if q.requires_grad), self._upd_params):
            out[name] = m
        return out

    def state_dict(self):
        """model.state_dict() with the fp32 master values for the parameters the modules hold as bf16 compute copies:
        what a checkpoint (DetectionCheckpointer, reference names and shapes) must contain."""
        sd = self.model.state_dict()
        for name, m in self.master_parameters().items():
            sd[name] = m.detach().clone()
        return sd

    def _momentum_by_name(self):
        """name -> momentum buffer (fp32) for every trainable parameter, whichever step implementation owns it."""
        names = [n for n, q in self.model.named_parameters() if q.requires_grad]
        if self.lowp:                                     # flat buffer with the masters' offsets
            out, base = {}, self._master_all.data_ptr()
            for n, m in zip(names, self._upd_params):
                off = (m.data_ptr() - base) // 4
                out[n] = _like_view(self._mom_all, off, m)
            return out
        if self.static_graph:                             # foreach step: per-group lists, created lazily
            if self._mom_bufs is None:
                self._mom_bufs = [[torch.zeros_like(p) for p in params] for params, _, _ in self._groups()]
            idx = {id(p): n for n, p in zip(names, self._upd_params)}
            out = {}
            for (params, _, _), bufs in zip(self._groups(), self._mom_bufs):
                for p, b in zip(params, bufs):
                    out[idx[id(p)]] = b
            return out
        out = {}
        for n, p in zip(names, self.params):              # eager step: torch.optim.SGD state
            st = self.optimizer.state.get(p, {})
            if st.get("momentum_buffer") is not None:
                out[n] = st["momentum_buffer"]
        return out

    def checkpoint(self):
        """What the reference's checkpointer saves (engine/train_loop.py + fvcore Checkpointer: model, optimizer,
        scheduler position, iteration): {"model": state_dict(), "optimizer": {"momentum": {name: buf}}, "iteration": it,
        "scaler": GradScaler state or None}. A resume from it is equivalent to never having stopped."""
        return {"model": self.state_dict(),
                "optimizer": {"momentum": {k: v.detach().clone() for k, v in self._momentum_by_name().items()}},
                "iteration": self.iter,
                "scaler": self.scaler.state_dict() if self.scaler is not None else None}

    @torch.no_grad()
    def load_checkpoint(self, ckpt, strict=True):
        res = self.load_state_dict(ckpt["model"], strict=strict)
        mom = ckpt.get("optimizer", {}).get("momentum", {})
        if mom:
            if not self.lowp and not self.static_graph:   # eager step: seed torch.optim.SGD's state
                names = [n for n, q in self.model.named_parameters() if q.requires_grad]
                for n, p in zip(names, self.params):
                    if n in mom:
                        self.optimizer.state[p]["momentum_buffer"] = mom[n].to(p.device, torch.float32).clone()
            else:
                own = self._momentum_by_name()
                for k, v in mom.items():
                    if k in own:
                        assert own[k].shape == v.shape, "momentum shape mismatch for %s" % k
                        own[k].copy_(v.to(own[k].device))
        self.iter = int(ckpt.get("iteration", self.iter))
        if self.scaler is not None and ckpt.get("scaler") is not None:
            self.scaler.load_state_dict(ckpt["scaler"])
        return res

    @torch.no_grad()
    def load_state_dict(self, sd, strict=True):
        """inverse of state_dict(): loads parameters (into the fp32 masters when they exist) and buffers. strict=False
        (what DetectionCheckpointer does for a backbone-only MODEL.WEIGHTS pretrain) skips absent / unknown keys and
        returns them: (missing_keys, unexpected_keys). Shapes must match exactly - no silent broadcasting."""
        masters = self.master_parameters()
        own = self.model.state_dict()
        missing = [k for k in own if k not in sd]
        unexpected = [k for k in sd if k not in own]
        if strict:
            assert not missing, "missing keys: %s" % missing[:5]
            assert not unexpected, "unexpected keys: %s" % unexpected[:5]
        for k, v in sd.items():
            dst = masters.get(k, own.get(k))
            if dst is None:
                continue
            if tuple(dst.shape) != tuple(v.shape):
                raise ValueError("shape mismatch for %s: checkpoint %s vs model %s" % (k, tuple(v.shape), tuple(dst.shape)))
            dst.copy_(v.to(dst.device))
        if self.lowp:
            self._w16_flat.copy_(self._master_flat)
        return missing, unexpected

    @torch.no_grad()
    def broadcast_parameters(self, src=0):
        """DDP's initial broadcast: rank `src`'s parameters and buffers everywhere."""
        if _world() == 1:
            return
        for t in list(self._upd_params) + list(self.model.buffers()):
            dist.broadcast(t, src)
        if self.lowp:
            self._w16_flat.copy_(self._master_flat)

    @torch.no_grad()
    def _fused_clip_sgd(self):
        """Per-parameter gradient-norm clipping + SGD(momentum, weight decay) + refresh of the bf16 compute weights in
        one pass over the flat buffers (csrc/optimizer.cu); the LR comes from a device scalar (graph-safe)."""
        from . import _lib
        import ctypes
        sv = self.cfg.SOLVER
        coef = None
        if self.clip is not None:
            assert self.clip.CLIP_TYPE == "norm"
            norms = torch.stack(torch._foreach_norm(self._seg_grads, self.clip.NORM_TYPE))
            coef = torch.clamp(self.clip.CLIP_VALUE / (norms + 1e-6), max=1.0).float().contiguous()
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None    # noqa: E731
        _lib.check(_lib.lib().u2b_sgd_step_segments(p(self.grads.flat), p(self._master_all), p(self._mom_all),
                                                    p(self._w16_flat), self.grads.n_head, p(self._seg_chunk),
                                                    p(self._seg_wd), p(coef), p(self._lr_t), float(sv.MOMENTUM),
                                                    int(bool(sv.NESTEROV)), self._master_all.numel(), _lib.stream_ptr()),
                   "u2b_sgd_step_segments")
        _lib.count_launches(1)

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
            if s.NESTEROV:                                 # torch.optim.SGD: d_p = grad + momentum * buf
                step = torch._foreach_add(grads, bufs, alpha=s.MOMENTUM)
                upd = torch._foreach_mul(step, self._lr_t)
            else:
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

    @torch.no_grad()
    def _gather_grads(self):
        """Every parameter's `.grad` (handed over by autograd; bf16 for the bf16 compute weights) -> its fp32 view in
        the flat gradient buffer, with as few multi-tensor copies as the layouts allow."""
        dst = self._upd_grads if self._upd_grads is not None else self._flat_views
        have = [(d, p.grad) for d, p in zip(dst, self.params) if p.grad is not None]
        if len(have) != len(self.params):
            self.grads.zero_()                # parameters outside the graph of this step keep a zero gradient
        # the multi-tensor fast path is all-or-nothing per call: keep same-dtype / same-layout pairs together
        buckets = {}
        for d, g in have:
            key = (g.dtype, d.stride() == g.stride())
            buckets.setdefault(key, ([], []))
            buckets[key][0].append(d)
            buckets[key][1].append(g)
        for dsts, srcs in buckets.values():
            torch._foreach_copy_(dsts, srcs)

    # ---- gradient all-reduce overlapped with the backward pass (data-parallel runs) ----
    def _setup_overlap(self):
        """The flat fp32 gradient buffer is cut into contiguous buckets (~U2B_BUCKET_MB each, whole parameters). A
        post-accumulate-grad hook per parameter counts its bucket down; when the last gradient of a bucket has been
        produced, the bucket is gathered into the flat buffer (widening bf16 weight gradients) and all-reduced (AVG) on
        a communication stream while autograd keeps going on the compute streams. Parameters are laid out in forward
        order, so the heads' buckets (half of all weights: the three 12544x1024 box-head FCs) finish first and their
        NVLink traffic hides behind the backbone's backward. Equivalent to the reference's DDP buckets
        (engine/defaults.py:60-79 create_ddp_model); captured into the step's CUDA graph like everything else."""
        dst = self._upd_grads if self._upd_grads is not None else self._flat_views
        base = self.grads.flat.data_ptr()
        order = sorted(range(len(self.params)), key=lambda i: dst[i].data_ptr())
        cap = int(float(os.environ.get("U2B_BUCKET_MB", "24")) * (1 << 20) / 4)
        buckets, cur, cur_n = [], [], 0
        for i in order:
            n = _aligned(self.params[i].numel())
            if cur and cur_n + n > cap:
                buckets.append(cur)
                cur, cur_n = [], 0
            cur.append(i)
            cur_n += n
        if cur:
            buckets.append(cur)
        on_gpu = self.device.type == "cuda"      # on the CPU (gloo tests of this logic) there are no streams: every bucket is
        self._ov = {"buckets": buckets, "dst": dst, "bucket_of": {}, "range": [],     # reduced in place when it completes
                    "comm": torch.cuda.Stream(self.device) if on_gpu else None}
        for b, idxs in enumerate(buckets):
            lo = (dst[idxs[0]].data_ptr() - base) // 4
            hi = (dst[idxs[-1]].data_ptr() - base) // 4 + _aligned(self.params[idxs[-1]].numel())
            self._ov["range"].append((lo, hi))
            for i in idxs:
                self._ov["bucket_of"][i] = b
        index_of = {id(p): i for i, p in enumerate(self.params)}

        def hook(p):
            ov = self._ov
            if not ov.get("armed"):
                return
            i = index_of[id(p)]
            b = ov["bucket_of"][i]
            if ov["comm"] is not None:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(self.device))      # the stream this gradient was produced on
                ov["events"][b].append(ev)
            ov["pending"][b] -= 1
            if ov["pending"][b] == 0:
                self._reduce_bucket(b)

        for p in self.params:
            p.register_post_accumulate_grad_hook(hook)

    @torch.no_grad()
    def _reduce_bucket(self, b, on_comm_stream=True):
        ov = self._ov
        idxs = ov["buckets"][b]
        lo, hi = ov["range"][b]
        import contextlib
        if ov["comm"] is None:
            ctx = contextlib.nullcontext()
        else:
            comm = ov["comm"] if on_comm_stream else torch.cuda.current_stream(self.device)
            if on_comm_stream:
                for ev in ov["events"][b]:
                    comm.wait_event(ev)
            ctx = torch.cuda.stream(comm)
        with ctx:
            have = [(ov["dst"][i], self.params[i].grad) for i in idxs if self.params[i].grad is not None]
            if len(have) != len(idxs):
                self.grads.flat[lo:hi].zero_()                     # parameters outside this step's graph: zero gradient
            groups = {}
            for d, g in have:
                groups.setdefault((g.dtype, d.stride() == g.stride()), ([], []))
                groups[(g.dtype, d.stride() == g.stride())][0].append(d)
                groups[(g.dtype, d.stride() == g.stride())][1].append(g)
            for dsts, srcs in groups.values():
                torch._foreach_copy_(dsts, srcs)
            _all_reduce_avg(self.grads.flat[lo:hi])
        ov["done"][b] = True

    def _arm_overlap(self):
        ov = self._ov
        n = len(ov["buckets"])
        ov["pending"] = [len(b) for b in ov["buckets"]]
        ov["events"] = [[] for _ in range(n)]
        ov["done"] = [False] * n
        if ov["comm"] is not None:
            ov["comm"].wait_stream(torch.cuda.current_stream(self.device))     # the flat buffer's previous consumers are done
        ov["armed"] = True
        # NCCL's all-reduce CTAs stay resident while backward runs: the persistent conv kernels leave them room (their grids
        # are fixed at capture time), otherwise the clusters that find no free SM pair run as a second wave
        budget = int(os.environ.get("U2B_OVERLAP_SM_BUDGET", "0"))
        if budget:
            from . import _lib
            _lib.check(_lib.lib().u2b_set_sm_budget(budget), "u2b_set_sm_budget")

    def _finish_overlap(self):
        ov = self._ov
        ov["armed"] = False
        if int(os.environ.get("U2B_OVERLAP_SM_BUDGET", "0")):
            from . import _lib
            _lib.check(_lib.lib().u2b_set_sm_budget(0), "u2b_set_sm_budget")
        if ov["comm"] is not None:
            torch.cuda.current_stream(self.device).wait_stream(ov["comm"])
        for b, done in enumerate(ov["done"]):                      # buckets with parameters that received no gradient
            if not done:
                self._reduce_bucket(b, on_comm_stream=False)

    def _static_step(self):
        from .modeling.static_train import forward_train_static
        # .grad = None: autograd hands over each gradient tensor instead of launching one `grad += g` kernel per
        # parameter; one multi-tensor copy then gathers (and, for bf16 weight gradients, widens) them into the flat
        # fp32 buffer. Under graph capture the gradient tensors live in the graph's private pool.
        for p in self.params:
            p.grad = None
        overlap = _world() > 1 and os.environ.get("U2B_OVERLAP_ALLREDUCE", "1") == "1"
        if overlap and getattr(self, "_ov", None) is None:
            self._setup_overlap()
        with torch.autocast("cuda", dtype=self.amp_dtype, enabled=self.amp_dtype is not None):
            loss_dict, flag = forward_train_static(self.model, *self._static_in)
        if overlap:
            self._arm_overlap()
            sum(loss_dict.values()).backward()
            self._finish_overlap()
        else:
            sum(loss_dict.values()).backward()
            self._gather_grads()
            self.grads.all_reduce_mean()      # one NCCL all-reduce of the flat gradient buffer (captured with the graph)
        if self.lowp and os.environ.get("U2B_FUSED_OPT", "1") != "0":
            self._fused_clip_sgd()
        else:
            self._clip_foreach()
            self._sgd_foreach()
        return {k: v.detach() for k, v in loss_dict.items()}, flag

    def _pack(self, batched_inputs):
        """Reference-format batch (host or device tensors) -> padded device tensors. Host tensors are copied one by
        one (pinned memory: asynchronous) before the device-side packing."""
        from .modeling.static_train import pack_batch
        dev_batch = []
        for d in batched_inputs:
            dev_batch.append({"image": d["image"].to(self.device, non_blocking=True),
                              "instances": d["instances"].to(self.device, non_blocking=True),
                              "sem_seg": d["sem_seg"].to(self.device, non_blocking=True)})
        g_max = self.g_max or max(1, max(len(d["instances"]) for d in batched_inputs))
        if self.g_max is None:
            self.g_max = g_max
        return pack_batch(dev_batch, self.device, g_max)

    def _load_static_inputs(self, batched_inputs):
        packed = self._pack(batched_inputs)
        if getattr(self, "_static_in", None) is None:
            self._static_in = [t.clone() for t in packed]
        else:
            for dst, src in zip(self._static_in, packed):
                assert dst.shape == src.shape, "static_graph needs constant input shapes (%s vs %s)" % (dst.shape, src.shape)
                dst.copy_(src, non_blocking=True)

    def prefetch(self, batched_inputs):
        """Stage the NEXT step's batch while the current step is still running on the GPU: host->device copies and
        the packing run on a copy stream into staging buffers; the following run_step(None) only copies them
        device-to-device into the graph's static inputs. (The reference overlaps data loading with compute through
        DataLoader workers, data/build.py; H2D itself happens in rcnn.py:223-234 preprocess_image.)"""
        assert self.static_graph, "prefetch() belongs to the static-graph step"
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._staged_evt, self._staging_free = torch.cuda.Event(), torch.cuda.Event()
            self._staging, self._staged = None, False
            self._staging_free.record(torch.cuda.current_stream(self.device))
        cs = self._copy_stream
        cs.wait_event(self._staging_free)          # the previous staged batch has been consumed
        with torch.cuda.stream(cs):
            packed = self._pack(batched_inputs)
            if self._staging is None:
                self._staging = [t.clone() for t in packed]
            else:
                for dst, src in zip(self._staging, packed):
                    assert dst.shape == src.shape, "static_graph needs constant input shapes"
                    dst.copy_(src, non_blocking=True)
            self._staged_evt.record(cs)
        self._staged = True

    def _consume_staged(self):
        main = torch.cuda.current_stream(self.device)
        main.wait_event(self._staged_evt)
        if getattr(self, "_static_in", None) is None:
            self._static_in = [t.clone() for t in self._staging]
        else:
            for dst, src in zip(self._static_in, self._staging):
                dst.copy_(src, non_blocking=True)
        self._staging_free.record(main)
        self._staged = False

    def _run_step_static(self, batched_inputs):
        self._lr_t.fill_(self.sched.lr(self.iter))
        if batched_inputs is None:
            assert getattr(self, "_staged", False), "run_step(None) needs a batch staged with prefetch()"
            self._consume_staged()
        else:
            self._load_static_inputs(batched_inputs)
        if self._graph is None:
            cur = torch.cuda.current_stream()
            side = torch.cuda.Stream()
            side.wait_stream(cur)
            # warm-up outside capture (lazy initialisations, autotuning, workspaces) must leave NO trace in the training
            # state: the reference applies one update per batch (train_loop.py:479-521). Parameters, momentum, BN
            # running statistics / counters and the device RNG are snapshotted and restored around it.
            snap = self._snapshot_training_state()
            with torch.cuda.stream(side):
                for _ in range(3):
                    self._static_step()
            cur.wait_stream(side)
            torch.cuda.synchronize()
            self._restore_training_state(snap)
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

    @torch.no_grad()
    def _snapshot_training_state(self):
        snap = {"buffers": [b.detach().clone() for b in self.model.buffers()],
                "rng": torch.cuda.get_rng_state(self.device), "cpu_rng": torch.get_rng_state()}
        if self.lowp:
            snap["flat"] = [t.clone() for t in (self._master_all, self._mom_all, self._w16_flat)]
        else:
            snap["params"] = [p.detach().clone() for p in self._upd_params]
        return snap

    @torch.no_grad()
    def _restore_training_state(self, snap):
        for b, v in zip(self.model.buffers(), snap["buffers"]):
            b.copy_(v)
        if self.lowp:
            for t, v in zip((self._master_all, self._mom_all, self._w16_flat), snap["flat"]):
                t.copy_(v)
        else:
            for p, v in zip(self._upd_params, snap["params"]):
                p.copy_(v)
            if self._mom_bufs is not None:          # created (zero) by the first foreach step
                for bufs in self._mom_bufs:
                    torch._foreach_zero_(bufs)
        torch.cuda.set_rng_state(snap["rng"], self.device)
        torch.set_rng_state(snap["cpu_rng"])

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
