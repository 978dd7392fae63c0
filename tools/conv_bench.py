"""Developer tool: tcgen05 conv (libu2b200) vs the library kernel (cuDNN) per hot-path shape, fwd only."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from u2seg_b200.modeling.conv_tc import conv2d_nhwc, set_cluster

SHAPES = [  # name, N, Cin, H, W, Cout, k, stride  (Appendix A, training)
    ("fpn_output2/rpn_p2", 2, 256, 256, 256, 256, 3, 1), ("fpn_output3", 2, 256, 128, 128, 256, 3, 1),
    ("semseg_p2", 2, 256, 256, 256, 128, 3, 1), ("mask_fcn", 256, 256, 14, 14, 256, 3, 1),
    ("res2_conv2", 2, 64, 256, 256, 64, 3, 1), ("res2_conv3", 2, 64, 256, 256, 256, 1, 1), ("res2_conv1", 2, 256, 256, 256, 64, 1, 1),
    ("res3_conv2", 2, 128, 128, 128, 128, 3, 1), ("res3_conv3", 2, 128, 128, 128, 512, 1, 1),
    ("res4_conv2", 2, 256, 64, 64, 256, 3, 1), ("res4_conv3", 2, 256, 64, 64, 1024, 1, 1), ("res4_conv1", 2, 1024, 64, 64, 256, 1, 1),
    ("res5_conv2", 2, 512, 32, 32, 512, 3, 1), ("res5_conv3", 2, 512, 32, 32, 2048, 1, 1),
    ("fpn_lateral2", 2, 256, 256, 256, 256, 1, 1), ("fc1_as_conv", 1, 12544, 1, 1024, 1024, 1, 1),
]
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
torch.backends.cudnn.benchmark = True
for name, N, Cin, H, W, Cout, k, s in SHAPES:
    x = torch.randn(N, Cin, H, W, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    if H == 1:
        x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    w = torch.randn(Cout, Cin, k, k, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
    wo = w.permute(0, 2, 3, 1).contiguous()
    pad = k // 2
    OH, OW = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    fl = 2.0 * N * OH * OW * Cout * Cin * k * k
    ts = []
    for cl in (1, 2, 4):
        set_cluster(cl)
        ts.append(timeit(lambda: conv2d_nhwc(x, wo, s, pad)))
    t2 = timeit(lambda: F.conv2d(x, w, None, s, pad))
    print("%-22s %8.1f GF | tc cl1 %.3f ms %6.1f TF/s | cl2 %.3f ms %6.1f | cl4 %.3f ms %6.1f | cudnn %.3f ms %6.1f TF/s"
          % (name, fl / 1e9, ts[0], fl / ts[0] / 1e9, ts[1], fl / ts[1] / 1e9, ts[2], fl / ts[2] / 1e9, t2, fl / t2 / 1e9))
