"""Training-mode SyncBatchNorm + residual add + ReLU as ONE autograd node over libu2b200's NHWC kernels
(csrc/batchnorm.cu). Semantics of nn.SyncBatchNorm (detectron2/layers/batch_norm.py:187): batch statistics over
the whole data-parallel group, running statistics updated with momentum (unbiased variance)."""
import ctypes

import torch
import torch.distributed as dist

from .. import _lib

_CODE = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}
_ws = {}


def _partials(S, C, device):
    """(S, 2C) fp32 workspace; one growing buffer per device (kernels are stream-ordered, each use is consumed by
    the finalize / coefficient kernel enqueued right after it)."""
    need = S * 2 * C
    key = (device, torch.cuda.current_stream(device).cuda_stream)     # one workspace per stream: no cross-stream reuse
    buf = _ws.get(key)
    if buf is None or buf.numel() < need:
        buf = torch.empty((max(need, 1 << 20),), dtype=torch.float32, device=device)
        _ws[key] = buf
    return buf


def _world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class _PeerExchange:
    """Symmetric (peer-mapped) buffers for the in-kernel SyncBN reduction over NVLink (csrc/batchnorm.cu
    xchg_all_reduce). Allocated once per process through torch's symmetric-memory allocator; every BN forward /
    backward of the step consumes one `epoch`, in the same order on every rank."""
    SLOT_FLOATS = 4096   # 2C for C <= 2048

    def __init__(self, device):
        import torch.distributed._symmetric_memory as symm
        L = _lib.lib()
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        nbytes = int(L.u2b_bn_xchg_buffer_bytes(self.world, self.SLOT_FLOATS))
        self.buf = symm.empty(nbytes // 4 + 16, dtype=torch.float32, device=device)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        self.hdl = symm.rendezvous(self.buf, dist.group.WORLD.group_name)
        self.peers = torch.tensor([int(p) for p in self.hdl.buffer_ptrs], dtype=torch.int64, device=device)
        self.epoch_ctr = torch.zeros((1,), dtype=torch.int32, device=device)   # advanced inside the kernels
        self.epoch_ctrs = torch.zeros((int(L.u2b_bn_xchg2_max_ctas()),), dtype=torch.int32, device=device)   # per-CTA sequences
        dist.barrier()


# 1: one launch per BN direction (multi-CTA, 8-byte {value, epoch} words); 0: round 1's bn_sum_partials + single-CTA exchange
XCHG_MULTI_CTA = __import__("os").environ.get("U2B_SYNCBN_XCHG2", "1") != "0"

_xchg = {"obj": None, "failed": False}


def peer_exchange(device):
    """The process-wide exchange object, or None (NCCL all-reduce path) when disabled / unavailable."""
    import os
    if _xchg["failed"] or os.environ.get("U2B_SYNCBN_XCHG", "1") == "0":
        return None
    if _xchg["obj"] is None:
        try:
            _xchg["obj"] = _PeerExchange(device)
        except Exception as e:   # no P2P / symmetric memory: keep NCCL
            import warnings
            warnings.warn("u2seg_b200: NVLink peer exchange for SyncBN unavailable (%r); using NCCL all-reduce" % (e,))
            _xchg["failed"] = True
            return None
    return _xchg["obj"]


def _nhwc(x):
    if x.is_contiguous(memory_format=torch.channels_last) and x.stride(1) == 1:
        return x
    return x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def channel_sum(x):
    """sum over (N,H,W) of a logical (N,C,H,W) tensor -> (C,) fp32: the bias gradient of a convolution. Runs on the
    BN statistics kernels (strip partial sums at HBM speed + one small column sum) instead of an fp32 copy of the
    tensor followed by a generic reduction."""
    L = _lib.lib()
    C = x.shape[1]
    if not (x.is_cuda and x.dtype in _CODE and x.dim() == 4 and x.numel() > 0 and L.u2b_bn_supported(C)):
        return x.sum(dim=(0, 2, 3), dtype=torch.float32)
    xc = _nhwc(x)
    P = xc.numel() // C
    s = _lib.stream_ptr()
    S = int(L.u2b_bn_num_strips(P, C))
    part = _partials(S, C, xc.device)
    _lib.check(L.u2b_bn_stats(_CODE[xc.dtype], _p(xc), P, C, _p(part), s), "u2b_bn_stats")
    sums = torch.empty((2 * C,), dtype=torch.float32, device=xc.device)
    _lib.check(L.u2b_bn_sum_partials(_p(part), S, 2 * C, _p(sums), s), "u2b_bn_sum_partials")
    _lib.count_launches(2)
    return sums[:C]


class _BNAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, running_mean, running_var, eps, momentum, relu, partials=None,
                residual_up2x=False):
        L = _lib.lib()
        s = _lib.stream_ptr()
        xc = _nhwc(x)
        N, C, H, W = xc.shape
        P = N * H * W
        dt = _CODE[xc.dtype]
        dev = xc.device
        world = _world()
        if partials is not None:      # (S, 2C) [sum | sumsq] rows from the producing conv's epilogue (csrc/conv2.cu)
            assert partials.shape[1] == 2 * C and partials.dtype == torch.float32 and partials.is_contiguous()
            part, S = partials, int(partials.shape[0])
        else:
            S = int(L.u2b_bn_num_strips(P, C))
            part = _partials(S, C, dev)
            _lib.check(L.u2b_bn_stats(dt, _p(xc), P, C, _p(part), s), "u2b_bn_stats")
        stats = torch.empty((4 * C,), dtype=torch.float32, device=dev)   # mean | invstd | scale | shift
        n_total = float(P) * world
        done = False
        if world > 1:
            px = peer_exchange(dev)
            if px is not None and 2 * C <= px.SLOT_FLOATS and not XCHG_MULTI_CTA:
                sums = torch.empty((2 * C,), dtype=torch.float32, device=dev)
                _lib.check(L.u2b_bn_sum_partials(_p(part), S, 2 * C, _p(sums), s), "u2b_bn_sum_partials")
                _lib.check(L.u2b_bn_xchg_finalize(_p(sums), _p(px.peers), px.world, px.rank, _p(px.epoch_ctr),
                                                  px.SLOT_FLOATS, n_total, _p(weight), _p(bias), float(eps),
                                                  float(momentum), _p(running_mean), _p(running_var), _p(stats), C, s),
                           "u2b_bn_xchg_finalize")
                done = True
            elif px is not None and 2 * C <= px.SLOT_FLOATS:
                # partial rows -> sums -> NVLink exchange -> statistics in ONE launch (one CTA per 32 channels)
                _lib.check(L.u2b_bn_xchg2_finalize(_p(part), S, _p(px.peers), px.world, px.rank, _p(px.epoch_ctrs),
                                                   px.SLOT_FLOATS, n_total, _p(weight), _p(bias), float(eps),
                                                   float(momentum), _p(running_mean), _p(running_var), _p(stats), C, s),
                           "u2b_bn_xchg2_finalize")
                done = True
            else:
                sums = torch.empty((2 * C,), dtype=torch.float32, device=dev)
                _lib.check(L.u2b_bn_sum_partials(_p(part), S, 2 * C, _p(sums), s), "u2b_bn_sum_partials")
                dist.all_reduce(sums)
                part, S = sums, 1
        if not done:
            _lib.check(L.u2b_bn_finalize(_p(part), S, n_total, _p(weight), _p(bias), float(eps), float(momentum),
                                         _p(running_mean), _p(running_var), _p(stats), C, s), "u2b_bn_finalize")
        res = _nhwc(residual.to(xc.dtype)) if residual is not None else None
        y = torch.empty((N, H, W, C), dtype=xc.dtype, device=dev).permute(0, 3, 1, 2)
        if residual_up2x:     # residual is the coarser pyramid level: added through a nearest x2 upsampling (fpn.py:153-156)
            assert res is not None and tuple(res.shape) == (N, C, H // 2, W // 2) and H % 2 == 0 and W % 2 == 0
            _lib.check(L.u2b_bn_apply_resup(dt, _p(xc), _p(stats), _p(res), int(relu), _p(y), N, H, W, C, s), "u2b_bn_apply_resup")
        else:
            _lib.check(L.u2b_bn_apply(dt, _p(xc), _p(stats), _p(res), int(relu), _p(y), P, C, s), "u2b_bn_apply")
        _lib.count_launches(3)
        # y = relu(bn(x)) without residual: the backward recomputes the mask from x and the saved scale / shift (exactly
        # y > 0) instead of reading y - one of three tensor reads less in both backward passes
        mask_from_x = relu and residual is None and RELU_MASK_FROM_X
        ctx.save_for_backward(xc, y if (relu and not mask_from_x) else torch.empty(0), weight, stats)
        ctx.meta = (relu, residual is not None, n_total, world, bool(residual_up2x), mask_from_x)
        return y

    @staticmethod
    def backward(ctx, gy):
        L = _lib.lib()
        s = _lib.stream_ptr()
        xc, y, weight, stats = ctx.saved_tensors
        relu, has_res, n_total, world, res_up, mask_from_x = ctx.meta
        N, C, H, W = xc.shape
        P = N * H * W
        dt = _CODE[xc.dtype]
        dev = xc.device
        g = _nhwc(gy.to(xc.dtype))
        yy = y if (relu and not mask_from_x) else None
        S = int(L.u2b_bn_num_strips(P, C))
        part = _partials(S, C, dev)
        if mask_from_x:
            _lib.check(L.u2b_bn_bwd_reduce_relu_x(dt, _p(g), _p(xc), _p(stats), P, C, _p(part), s), "u2b_bn_bwd_reduce_relu_x")
        else:
            _lib.check(L.u2b_bn_bwd_reduce(dt, _p(g), _p(xc), _p(yy), _p(stats), P, C, _p(part), s), "u2b_bn_bwd_reduce")
        coeff = torch.empty((3 * C,), dtype=torch.float32, device=dev)
        gwb = torch.empty((2 * C,), dtype=torch.float32, device=dev)      # dgamma | dbeta (LOCAL sums: DDP reduces them)
        if world > 1:
            px = peer_exchange(dev)
            if px is not None and 2 * C <= px.SLOT_FLOATS and not XCHG_MULTI_CTA:
                sums = torch.empty((2 * C,), dtype=torch.float32, device=dev)
                _lib.check(L.u2b_bn_sum_partials(_p(part), S, 2 * C, _p(sums), s), "u2b_bn_sum_partials")
                _lib.check(L.u2b_bn_xchg_bwd_coeff(_p(sums), _p(px.peers), px.world, px.rank, _p(px.epoch_ctr),
                                                   px.SLOT_FLOATS, n_total, _p(stats), _p(weight), _p(coeff), _p(gwb),
                                                   C, s), "u2b_bn_xchg_bwd_coeff")
            elif px is not None and 2 * C <= px.SLOT_FLOATS:
                _lib.check(L.u2b_bn_xchg2_bwd_coeff(_p(part), S, _p(px.peers), px.world, px.rank, _p(px.epoch_ctrs),
                                                    px.SLOT_FLOATS, n_total, _p(stats), _p(weight), _p(coeff), _p(gwb),
                                                    C, s), "u2b_bn_xchg2_bwd_coeff")
            else:
                sums = torch.empty((2 * C,), dtype=torch.float32, device=dev)
                _lib.check(L.u2b_bn_sum_partials(_p(part), S, 2 * C, _p(sums), s), "u2b_bn_sum_partials")
                gwb[:C].copy_(sums[C:])
                gwb[C:].copy_(sums[:C])
                dist.all_reduce(sums)
                _lib.check(L.u2b_bn_bwd_coeff(_p(sums), 1, n_total, _p(stats), _p(weight), _p(coeff), None, C, s),
                           "u2b_bn_bwd_coeff")
        else:
            _lib.check(L.u2b_bn_bwd_coeff(_p(part), S, n_total, _p(stats), _p(weight), _p(coeff), _p(gwb), C, s), "u2b_bn_bwd_coeff")
        dx = torch.empty((N, H, W, C), dtype=xc.dtype, device=dev).permute(0, 3, 1, 2)
        # without ReLU the residual's gradient IS the incoming gradient (no masked copy needed)
        need_dres = has_res and (relu or not res_up)
        dres = torch.empty((N, H, W, C), dtype=xc.dtype, device=dev).permute(0, 3, 1, 2) if need_dres else None
        if mask_from_x:
            _lib.check(L.u2b_bn_bwd_apply_relu_x(dt, _p(g), _p(xc), _p(stats), _p(coeff), _p(dx), P, C, s), "u2b_bn_bwd_apply_relu_x")
        else:
            _lib.check(L.u2b_bn_bwd_apply(dt, _p(g), _p(xc), _p(yy), _p(coeff), _p(dx), _p(dres), P, C, s), "u2b_bn_bwd_apply")
        _lib.count_launches(3)
        if has_res and res_up:      # gradient of the nearest x2 upsampling: fold every 2x2 block (csrc/pool.cu)
            src = dres if dres is not None else g
            if dt in (1, 2):
                folded = torch.empty((N, H // 2, W // 2, C), dtype=xc.dtype, device=dev).permute(0, 3, 1, 2)
                _lib.check(L.u2b_sum2x2_nhwc(dt, _p(src), N, H, W, C, _p(folded), s), "u2b_sum2x2_nhwc")
                _lib.count_launches(1)
            else:
                folded = src.reshape(N, C, H // 2, 2, W // 2, 2).sum(dim=(3, 5))
            dres = folded
        return dx, gwb[:C].to(weight.dtype), gwb[C:].to(weight.dtype), dres, None, None, None, None, None, None, None


def bn_act(x, bn, residual=None, relu=False, partials=None, residual_up2x=False):
    """relu(SyncBN_train(x) + residual) for a BatchNorm2d-like module `bn` in training mode. `partials`: optional
    per-tile statistics of x computed by the kernel that produced it (skips the reduction pass). residual_up2x: the
    residual is the next-coarser map, added through a nearest-neighbour x2 upsampling (FPN top-down path)."""
    return _BNAct.apply(x, bn.weight, bn.bias, residual, bn.running_mean, bn.running_var, bn.eps,
                        0.1 if bn.momentum is None else bn.momentum, relu, partials, residual_up2x)


# The fused multi-GPU path takes the group's element count as P * world, i.e. it assumes every rank holds the same
# N*H*W. That is guaranteed only by the static-graph trainer (fixed input shapes; engine.Trainer sets this flag). The
# reference-shaped eager step pads each rank's multi-scale batch on its own (MIN_SIZE_TRAIN 240..1024), so there the
# model falls back to nn.SyncBatchNorm's own function, which all-gathers the per-rank counts (ops.batch_norm).
EQUAL_SHAPES_ACROSS_RANKS = False
RELU_MASK_FROM_X = __import__("os").environ.get("U2B_BN_MASK_FROM_X", "1") == "1"


def supported(x, bn):
    if _world() > 1 and not EQUAL_SHAPES_ACROSS_RANKS:
        return False
    return (x.is_cuda and x.dtype in _CODE and x.dim() == 4 and bn.training and bn.affine
            and bool(_lib.lib().u2b_bn_supported(int(x.shape[1]))))


class _GNAct(torch.autograd.Function):
    """relu(GroupNorm(x)) on NHWC activations (layers/batch_norm.py get_norm("GN") = nn.GroupNorm(32, C), used by
    the semantic head, semantic_seg.py:180-200). The reference's kernel needs NCHW-contiguous fp32: under channels_last
    autocast that is a layout copy in, a layout copy out and the same again in backward. Here each image goes through
    the strip-reduce / apply kernels of csrc/batchnorm.cu in place (NHWC, activation dtype) with the group statistics
    folded by u2b_gn_finalize / u2b_gn_bwd_coeff; ReLU is fused in both directions."""

    @staticmethod
    def forward(ctx, x, weight, bias, G, eps, relu):
        L = _lib.lib()
        s = _lib.stream_ptr()
        xc = _nhwc(x)
        N, C, H, W = xc.shape
        HW = H * W
        dt = _CODE[xc.dtype]
        dev = xc.device
        es = xc.element_size()
        w32, b32 = weight.detach().float().contiguous(), bias.detach().float().contiguous()
        S = int(L.u2b_bn_num_strips(HW, C))
        part = _partials(S, C, dev)
        stats = torch.empty((N, 4 * C), dtype=torch.float32, device=dev)
        y = torch.empty((N, H, W, C), dtype=xc.dtype, device=dev).permute(0, 3, 1, 2)
        for n in range(N):
            xn = ctypes.c_void_p(xc.data_ptr() + n * HW * C * es)
            yn = ctypes.c_void_p(y.data_ptr() + n * HW * C * es)
            st = ctypes.c_void_p(stats.data_ptr() + n * 4 * C * 4)
            _lib.check(L.u2b_bn_stats(dt, xn, HW, C, _p(part), s), "u2b_bn_stats")
            _lib.check(L.u2b_gn_finalize(_p(part), S, HW, G, _p(w32), _p(b32), float(eps), st, C, s), "u2b_gn_finalize")
            _lib.check(L.u2b_bn_apply(dt, xn, st, None, int(relu), yn, HW, C, s), "u2b_bn_apply")
        _lib.count_launches(3 * N)
        ctx.save_for_backward(xc, y if relu else torch.empty(0), w32, stats)
        ctx.meta = (G, relu, weight.dtype, bias.dtype)
        return y

    @staticmethod
    def backward(ctx, gy):
        L = _lib.lib()
        s = _lib.stream_ptr()
        xc, y, w32, stats = ctx.saved_tensors
        G, relu, wdt, bdt = ctx.meta
        N, C, H, W = xc.shape
        HW = H * W
        dt = _CODE[xc.dtype]
        dev = xc.device
        es = xc.element_size()
        g = _nhwc(gy.to(xc.dtype))
        S = int(L.u2b_bn_num_strips(HW, C))
        part = _partials(S, C, dev)
        coeff = torch.empty((3 * C,), dtype=torch.float32, device=dev)
        gwb = torch.empty((2 * C,), dtype=torch.float32, device=dev)
        dx = torch.empty((N, H, W, C), dtype=xc.dtype, device=dev).permute(0, 3, 1, 2)
        for n in range(N):
            off = n * HW * C * es
            xn, gn, dn = (ctypes.c_void_p(t.data_ptr() + off) for t in (xc, g, dx))
            yn = ctypes.c_void_p(y.data_ptr() + off) if relu else None
            st = ctypes.c_void_p(stats.data_ptr() + n * 4 * C * 4)
            _lib.check(L.u2b_bn_bwd_reduce(dt, gn, xn, yn, st, HW, C, _p(part), s), "u2b_bn_bwd_reduce")
            _lib.check(L.u2b_gn_bwd_coeff(_p(part), S, HW, G, st, _p(w32), _p(coeff), _p(gwb), int(n > 0), C, s),
                       "u2b_gn_bwd_coeff")
            _lib.check(L.u2b_bn_bwd_apply(dt, gn, xn, yn, _p(coeff), dn, None, HW, C, s), "u2b_bn_bwd_apply")
        _lib.count_launches(3 * N)
        return dx, gwb[:C].to(wdt), gwb[C:].to(bdt), None, None, None


def gn_supported(x, gn):
    return (x.is_cuda and x.dtype in _CODE and x.dim() == 4 and gn.affine and x.numel() > 0
            and bool(_lib.lib().u2b_gn_supported(int(x.shape[1]), int(gn.num_groups))))


def gn_act(x, gn, relu=False):
    """relu(GroupNorm(x)) for an nn.GroupNorm module; output keeps x's dtype and NHWC storage."""
    return _GNAct.apply(x, gn.weight, gn.bias, int(gn.num_groups), float(gn.eps), bool(relu))
