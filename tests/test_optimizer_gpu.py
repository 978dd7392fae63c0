"""GPU parity of the fused clip + SGD + bf16-refresh kernel (csrc/optimizer.cu) against torch.optim.SGD with
per-parameter clip_grad_norm_ (detectron2/solver/build.py:63-73,119-139)."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu


def _al(n):
    return (n + 63) // 64 * 64


@pytest.mark.parametrize("nesterov", [False, True])
def test_fused_clip_sgd_matches_torch_sgd(nesterov):
    from u2seg_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(3)
    shapes = [(100,), (64, 3, 3, 3), (1000, 17), (5,)]
    wds = [0.0, 1e-4, 1e-4, 0.0]
    params = [torch.randn(s, generator=g).cuda() for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.SGD([{"params": [r], "weight_decay": w} for r, w in zip(ref, wds)], lr=0.02, momentum=0.9,
                          nesterov=nesterov)
    n = sum(_al(p.numel()) for p in params)
    n_head = _al(params[0].numel())                       # first parameter stays fp32-only, the rest get bf16 copies
    master, mom, grad = (torch.zeros(n, device="cuda") for _ in range(3))
    w16 = torch.zeros(n - n_head, dtype=torch.bfloat16, device="cuda")
    chunk = torch.full((n // 64,), -1, dtype=torch.int32)
    offs, off = [], 0
    for i, p in enumerate(params):
        master[off:off + p.numel()] = p.flatten()
        chunk[off // 64:(off + _al(p.numel())) // 64] = i
        offs.append(off)
        off += _al(p.numel())
    chunk = chunk.cuda()
    seg_wd = torch.tensor(wds, dtype=torch.float32, device="cuda")
    lr = torch.tensor(0.02, dtype=torch.float32, device="cuda")
    pp = lambda t: ctypes.c_void_p(t.data_ptr())    # noqa: E731
    for step in range(3):
        grads = [torch.randn(s, generator=g).cuda() * (10.0 if step == 1 else 0.1) for s in shapes]   # step 1 clips
        for r, gr, o in zip(ref, grads, offs):
            r.grad = gr.clone()
            torch.nn.utils.clip_grad_norm_(r, 1.0, 2.0)
            grad[o:o + gr.numel()] = gr.flatten()
        opt.step()
        norms = torch.stack([gr.norm(2.0) for gr in grads])
        coef = torch.clamp(1.0 / (norms + 1e-6), max=1.0).float().contiguous()
        _lib.check(L.u2b_sgd_step_segments(pp(grad), pp(master), pp(mom), pp(w16), n_head, pp(chunk), pp(seg_wd), pp(coef),
                                           pp(lr), 0.9, int(nesterov), n, _lib.stream_ptr()), "u2b_sgd_step_segments")
        for r, o in zip(ref, offs):
            got = master[o:o + r.numel()].view_as(r)
            assert torch.allclose(got, r.detach(), rtol=2e-6, atol=1e-7), (step, o)          # FLOAT: fma vs mul+add
        assert torch.equal(w16, master[n_head:].bfloat16())                                      # bf16 refresh: exact rounding
    pad = torch.ones(n, dtype=torch.bool)
    for r, o in zip(ref, offs):
        pad[o:o + r.numel()] = False
    assert float(master[pad.cuda()].abs().max()) == 0.0       # padding untouched (zero grads, zero params)
