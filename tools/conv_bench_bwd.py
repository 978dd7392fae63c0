"""Developer tool: backward kernels per hot-path shape (bf16): input gradient on conv2 (forward filter read MN-major) and
weight gradient on conv_wgrad2 (2-CTA tcgen05) vs the library (cuDNN convolution_backward). L2 flushed between launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2seg_b200.modeling.conv_tc import conv2_nhwc_dgrad, conv_wgrad2, wgrad2_supported, conv2_dgrad_supported
from tools.conv_bench2 import SHAPES  # noqa
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / n
torch.backends.cudnn.benchmark = True
only = sys.argv[1:]
for name, N, Cin, H, W, Cout, k, s in SHAPES:
    if only and not any(o in name for o in only): continue
    pad = k // 2
    OH, OW = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    x = torch.randn(N, Cin, H, W, device="cuda").bfloat16().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    gy = torch.randn(N, Cout, OH, OW, device="cuda").bfloat16().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    w = torch.randn(Cout, Cin, k, k, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    wo = w.permute(0, 2, 3, 1).contiguous()
    fl = 2.0 * N * OH * OW * Cout * Cin * k * k
    tf = lambda t: fl / t / 1e9 if t == t else float("nan")
    td = tw = float("nan")
    if s == 1 and conv2_dgrad_supported(gy, w, s, pad):
        td = timeit(lambda: conv2_nhwc_dgrad(gy, wo, pad))
    if wgrad2_supported(x, Cout, k, k, s, pad):
        tw = timeit(lambda: conv_wgrad2(x, gy, k, k, s, pad, torch.bfloat16))
    cd = timeit(lambda: torch.ops.aten.convolution_backward(gy, x, w, None, [s, s], [pad, pad], [1, 1], False, [0, 0], 1, [True, False, False]))
    cw = timeit(lambda: torch.ops.aten.convolution_backward(gy, x, w, None, [s, s], [pad, pad], [1, 1], False, [0, 0], 1, [False, True, False]))
    print("%-20s %7.1f GF | dgrad conv2 %.3f ms %6.0f TF/s vs cudnn %.3f ms %6.0f (%.2fx) | wgrad2 %.3f ms %6.0f TF/s vs cudnn %.3f ms %6.0f (%.2fx)"
          % (name, fl / 1e9, td, tf(td), cd, tf(cd), cd / td, tw, tf(tw), cw, tf(cw), cw / tw), flush=True)
