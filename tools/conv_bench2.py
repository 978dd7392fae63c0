"""Developer tool: conv_tc (1-CTA, cluster 2) vs conv2 (2-CTA UMMA, +/- fused BN statistics) vs the library kernel
(cuDNN) per hot-path shape (training), forward only, bf16. L2 is flushed between timed launches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from u2seg_b200.modeling.conv_tc import conv2d_nhwc, conv2_nhwc, set_cluster, set_tile_n

SHAPES = [  # name, N, Cin, H, W, Cout, k, stride
    ("fpn_output2/rpn_p2", 2, 256, 256, 256, 256, 3, 1), ("fpn_output3", 2, 256, 128, 128, 256, 3, 1),
    ("semseg_p2", 2, 256, 256, 256, 128, 3, 1), ("mask_fcn", 256, 256, 14, 14, 256, 3, 1),
    ("res2_conv2", 2, 64, 256, 256, 64, 3, 1), ("res2_conv3", 2, 64, 256, 256, 256, 1, 1), ("res2_conv1", 2, 256, 256, 256, 64, 1, 1),
    ("res3_conv2", 2, 128, 128, 128, 128, 3, 1), ("res3_conv3", 2, 128, 128, 128, 512, 1, 1), ("res3_conv1", 2, 512, 128, 128, 128, 1, 1),
    ("res3.0_conv2_s2", 2, 128, 256, 256, 128, 3, 2), ("res3.0_shortcut_s2", 2, 256, 256, 256, 512, 1, 2),
    ("res4_conv2", 2, 256, 64, 64, 256, 3, 1), ("res4_conv3", 2, 256, 64, 64, 1024, 1, 1), ("res4_conv1", 2, 1024, 64, 64, 256, 1, 1),
    ("res5_conv2", 2, 512, 32, 32, 512, 3, 1), ("res5_conv3", 2, 512, 32, 32, 2048, 1, 1), ("res5_conv1", 2, 2048, 32, 32, 512, 1, 1),
    ("fpn_lateral2", 2, 256, 256, 256, 256, 1, 1), ("fpn_lateral5", 2, 2048, 32, 32, 256, 1, 1),
    ("fc1_as_conv", 1, 12544, 1, 1024, 1024, 1, 1), ("fc2_as_conv", 1, 1024, 1, 1024, 1024, 1, 1),
    ("mask_deconv_as_1x1", 256, 256, 14, 14, 1024, 1, 1),
]
if __name__ == "__main__":
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
        x = torch.randn(N, Cin, H, W, device="cuda").bfloat16().permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        w = torch.randn(Cout, Cin, k, k, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last)
        wo = w.permute(0, 2, 3, 1).contiguous()
        pad = k // 2
        OH, OW = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
        fl = 2.0 * N * OH * OW * Cout * Cin * k * k
        set_cluster(2)
        t1 = timeit(lambda: conv2d_nhwc(x, wo, s, pad))
        res = []
        for bn in (0, 256, 128, 64):
            if bn and Cout % bn: res.append(float("nan")); continue
            set_tile_n(bn)
            res.append(timeit(lambda: conv2_nhwc(x, wo, s, pad)))
        set_tile_n(0)
        ts = timeit(lambda: conv2_nhwc(x, wo, s, pad, want_stats=True))
        t2 = timeit(lambda: F.conv2d(x, w, None, s, pad))
        tf = lambda t: fl / t / 1e9
        print("%-20s %7.1f GF | tc1 %.3f ms %6.0f | conv2 auto %.3f ms %6.0f TF/s (bn256 %.3f bn128 %.3f bn64 %.3f) +stats %.3f | cudnn %.3f ms %6.0f TF/s | conv2/cudnn %.2fx"
              % (name, fl / 1e9, t1, tf(t1), res[0], tf(res[0]), res[1], res[2], res[3], ts, t2, tf(t2), t2 / res[0]), flush=True)
