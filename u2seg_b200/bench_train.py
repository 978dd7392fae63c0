"""bench.py workload `train`: u2seg_R50_800 training step, batch 2 / GPU, synthetic 1024x1024
COCO-panoptic-shaped inputs (BASELINE.json configs[1]; configs[2] under torchrun)."""
import json
import os
import time

import torch
import torch.distributed as dist

IMS_PER_GPU, H, W, NUM_CLASSES, SEM_CLASSES = 2, 1024, 1024, 800, 28
# analytic conv/GEMM work of one step (2 images): 673.8 GMAC fwd -> x2 flop x3 (fwd+dgrad+wgrad) = 2.021
# TFLOP/image with the reference's all-class mask predictor. This build computes only the selected class
# (mask_head.forward_selected), so the 41.1 GMAC of the 800-channel predictor are NOT counted: 3.797 TFLOP / step.
FLOP_PER_IMAGE = 1.8985e12


def _to_device(batch, dev):
    out = []
    for d in batch:
        out.append({"image": d["image"].to(dev, non_blocking=True), "instances": d["instances"].to(dev, non_blocking=True),
                    "sem_seg": d["sem_seg"].to(dev, non_blocking=True), "height": d["height"], "width": d["width"]})
    return out


def _batch_bytes(batch):
    n = 0
    for d in batch:
        n += d["image"].numel() * d["image"].element_size() + d["sem_seg"].numel() * d["sem_seg"].element_size()
        i = d["instances"]
        n += i.gt_boxes.tensor.numel() * 4 + i.gt_classes.numel() * 8 + i.gt_masks.tensor.numel()
    return n


def run_train(args, ClockSampler, load_peaks, dist_info, run_kmeans=None):
    from . import _lib
    from .config import get_u2seg_cfg
    from .data_synth import synthetic_batch
    from .engine import Trainer

    rank, world, local = dist_info()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = load_peaks()
    torch.backends.cudnn.benchmark = False   # ROI counts vary per step: autotuning every new shape costs far more than it saves
    cfg = get_u2seg_cfg(NUM_CLASSES)
    torch.manual_seed(0)
    static = os.environ.get("U2B_STATIC_GRAPH", "1") != "0"
    trainer = Trainer(cfg, amp_dtype=torch.bfloat16, device=dev, static_graph=static, g_max=20)
    trainer.broadcast_parameters(0)   # identical initial weights on every rank (DDP broadcasts rank 0's)

    pool_n = 4
    host_pool = [synthetic_batch(IMS_PER_GPU, H, W, NUM_CLASSES, SEM_CLASSES, seed=1234 + 97 * rank + i, pin=True)
                 for i in range(pool_n)]
    dev_pool = [_to_device(b, dev) for b in host_pool]
    warm = max(3, args.warmup)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()        # ranks finish building their (CPU-generated) batch pools at different times
    for i in range(warm):
        trainer.run_step(dev_pool[i % pool_n])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    # ---- value: inputs resident in HBM ----
    sampler = ClockSampler(local)
    l0 = _lib.launch_count
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    last = None
    for i in range(args.steps):
        last = trainer.run_step(dev_pool[i % pool_n])
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = _lib.launch_count - l0
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    value = IMS_PER_GPU * world / (ms_step * 1e-3)
    loss_total = float(sum(last.values()))

    # ---- e2e: public API with HOST (pinned) inputs; H2D every step, loss read back every step ----
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(3, min(args.steps, 20))
    d2h = 0

    pinned = torch.empty((2, 16), dtype=torch.float32).pin_memory()
    read_evt = [torch.cuda.Event(), torch.cuda.Event()]
    host_hist = []

    def e2e_loop(n):
        """Every step: the batch crosses PCIe (pinned host -> device, on a copy stream while the previous step computes) and
        the step's 10 losses are copied to pinned host memory and read. The read of step i happens after step i+1 has been
        launched (asynchronous logging), so the device never waits for the host between steps."""
        nonlocal d2h
        if static:
            trainer.prefetch(host_pool[0])
            for i in range(n):
                losses = trainer.run_step(None)
                dev_losses = torch.stack(list(losses.values())).float()
                k = dev_losses.numel()
                pinned[i % 2, :k].copy_(dev_losses, non_blocking=True)     # D2H of this step's result
                read_evt[i % 2].record()
                d2h = k * 4
                if i + 1 < n:
                    trainer.prefetch(host_pool[(i + 1) % pool_n])
                if i > 0:
                    read_evt[(i - 1) % 2].synchronize()
                    host_hist.append(pinned[(i - 1) % 2, :k].sum().item())
            read_evt[(n - 1) % 2].synchronize()
            host_hist.append(pinned[(n - 1) % 2, :k].sum().item())
        else:
            for i in range(n):
                losses = trainer.run_step(_to_device(host_pool[i % pool_n], dev))
                host_losses = torch.stack(list(losses.values())).float().cpu()    # D2H of the step's result
                d2h = host_losses.numel() * 4
                host_hist.append(float(host_losses.sum()))

    e2e_loop(3)      # untimed: first use of the copy stream, staging buffers and the read-back kernels (lazy module loading)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_loop(e2e_steps)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / e2e_steps
    tt = torch.tensor([dt], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = IMS_PER_GPU * world / float(tt)

    achieved = FLOP_PER_IMAGE * (value / world) / 1e12
    conv_roof = in_step_conv_roofline(trainer, dev_pool[0], peaks, ms_step)
    iso_roof = conv_tc_roofline(peaks)
    line = {
        "metric": "u2seg_R50_800_train_images_per_sec", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "u2seg_R50_800.yaml training step (PanopticFPN R50-FPN, SyncBN, CascadeROIHeads, "
                               "800 classes), batch 2/GPU, synthetic 1024x1024 COCO-panoptic-shaped inputs, "
                               "fwd+bwd+allreduce+clip+SGD", "global_batch": IMS_PER_GPU * world,
                   "image_size": [H, W], "parallelism": "dp%d" % world,
                   "l2_note": "activations per step (>1.8 GB) exceed the 126 MB L2; a pool of %d distinct batches is cycled" % pool_n},
        "clocks": clocks, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": _batch_bytes(host_pool[0]),
                "d2h_bytes_per_step": d2h,
                "what": "Trainer.prefetch(host batch) + Trainer.run_step(): pinned host inputs (uint8 images, bit masks, "
                        "sem_seg) copied H2D on a copy stream while the previous step computes; the 10 losses of every "
                        "step copied to pinned host memory and read one step later (asynchronous logging)",
                "steps": e2e_steps, "last_host_loss_total": host_hist[-1] if host_hist else None},
        "roofline": conv_roof,
        "roofline_isolated": iso_roof,
        "step_roofline": {"bound": "tensor", "what": "whole training step: conv/GEMM flop of SURVEY 8(d) "
                                                     "minus the unused mask-predictor channels (1.8985 TFLOP/image fwd+bwd) / step time",
                          "achieved": achieved, "peak": peaks["tf_sus"], "unit": "TFLOP/s",
                          "frac": achieved / peaks["tf_sus"], "peak_source": peaks["src"] + " bf16 sustained"},
        "conv_policy": _conv_policy_text(),
        "step_mode": ("static shapes (fixed-capacity device buffers, no host sync), forward+backward+all-reduce+clip+SGD "
                      "replayed from one CUDA graph") if static else "dynamic shapes (reference-shaped), eager",
        "final_loss": loss_total,
    }
    if run_kmeans is not None and not os.environ.get("U2B_BENCH_SKIP_KMEANS"):
        km = run_kmeans(args, emit=False)     # (the trainer and its graph stay alive: 180 GB of HBM is ample)     # second half of BASELINE.json's metric: k-means embeddings/s
        line["kmeans"] = {k: km[k] for k in ("metric", "value", "unit", "ms_per_step", "dtype", "roofline", "e2e",
                                             "gpu_launches", "clocks", "config", "scaling", "cpu_baseline") if k in km}
    if rank == 0 and world == 1 and not os.environ.get("U2B_BENCH_SKIP_INFER"):
        # BASELINE.json configs[4]: u2seg_R50_300 panoptic inference + the ROIAlign / paste_masks HBM rooflines
        from .bench_infer import run_infer
        del trainer
        torch.cuda.empty_cache()
        inf = run_infer(args, ClockSampler, load_peaks, dist_info, emit=False)
        line["infer"] = {k: inf[k] for k in ("metric", "value", "unit", "ms_per_step", "dtype", "e2e", "gpu_launches", "clocks",
                                            "config", "roofline", "rooflines") if k in inf}
    if rank == 0 and world == 1 and not os.environ.get("U2B_BENCH_SKIP_CPU"):
        line["cpu_baseline"] = cpu_train_sample(1)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # A process group whose collectives were captured in a CUDA graph can block in its destructor; everything
        # is measured and printed, so leave without running destructors (all ranks, after a final barrier).
        dist.barrier()
        torch.cuda.synchronize()
        import sys
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def _conv_policy_text():
    from .modeling import conv_tc, ops
    if ops.TCGEN05_CONV_POLICY == "all":
        return ("tcgen05 2-CTA kernels (csrc/conv2.cu, conv_wgrad2.cu) for every conv with Cin,Cout %% 64 == 0 (1x1 and 3x3, "
                "stride 1 and 2: forward; stride-1 input gradient; weight gradient when Cout or Cin %% 256 == 0 and the layer has "
                ">= %g GFLOP), the box-head Linear layers (forward, input gradient) and the mask head's 2x2 transposed convolution; "
                "library (cuDNN/cuBLAS) for the rest (Cout 3/12/28 heads, smaller weight gradients, stride-2 input gradients, "
                "FC weight gradients); conv2=%s wgrad2=%s" % (conv_tc.WGRAD2_MIN_GFLOP, conv_tc.USE_CONV2, conv_tc.USE_WGRAD2))
    return "policy %s (conv2=%s wgrad2=%s)" % (ops.TCGEN05_CONV_POLICY, conv_tc.USE_CONV2, conv_tc.USE_WGRAD2)


def in_step_conv_roofline(trainer, batch, peaks, ms_step):
    """The dominant hand-written kernel AS IT RUNS INSIDE THE STEP: one extra, eager execution of the static step (the
    same kernel sequence the CUDA graph replays: same shapes, same tensors, cache state of a real step) with every
    tcgen05 launch bracketed by CUDA events on its launching stream. Launches are grouped by (direction, shape); the
    group with the most time is reported against the SUSTAINED bf16 peak (a kernel timed inside a long step), next to
    the aggregate over all tcgen05 launches of the step. Training state is restored afterwards."""
    from .modeling import conv_tc
    if not trainer.static_graph:
        return conv_tc_roofline(peaks)
    snap = trainer._snapshot_training_state()
    trainer._load_static_inputs(batch)
    torch.cuda.synchronize()
    from .modeling import static_train
    static_train.FORCE_SINGLE_STREAM = True     # no concurrent branch kernels: the events bracket each kernel alone
    how = "graph"
    try:
        try:
            # the step captured ONCE MORE as a single-stream CUDA graph whose tcgen05 launches are bracketed by external
            # event-record nodes: replayed, the events time each kernel on the device with no host launch gap in between
            conv_tc.TIMING, conv_tc.TIMING_EXTERNAL = [], True
            mg = torch.cuda.CUDAGraph()
            with torch.cuda.graph(mg):
                trainer._static_step()
            recs = conv_tc.TIMING
            conv_tc.TIMING = None
            for _ in range(2):
                mg.replay()
            torch.cuda.synchronize()
            _ = recs[0][3].elapsed_time(recs[0][4])
        except Exception as e:      # external events unavailable: eager step behind a long device-side sleep
            how = "eager (%s)" % type(e).__name__
            conv_tc.TIMING, conv_tc.TIMING_EXTERNAL = [], False
            torch.cuda._sleep(int(2e8))
            trainer._static_step()
            torch.cuda.synchronize()
            recs = conv_tc.TIMING
    finally:
        conv_tc.TIMING, conv_tc.TIMING_EXTERNAL = None, False
        static_train.FORCE_SINGLE_STREAM = False
    trainer._restore_training_state(snap)
    groups = {}
    for kind, key, flop, e0, e1 in recs:
        gkey = (kind, key)
        d = groups.setdefault(gkey, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += e0.elapsed_time(e1)
        d[2] += flop
    if not groups:
        return conv_tc_roofline(peaks)
    tot_ms = sum(v[1] for v in groups.values())
    tot_flop = sum(v[2] for v in groups.values())
    (kind, key), (cnt, ms, flop) = max(groups.items(), key=lambda kv: kv[1][1])
    N, Hh, Ww, Cin, Cout, R, stride = key
    # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures (cold L2)
    ncu = {("wgrad", (256, 14, 14, 256, 256, 3, 1)): (51.45e6 + 2.90e6, "profiles/r02_ncu_wgrad2_mask_head_summary.txt"),
           ("fwd", (2, 256, 256, 256, 256, 3, 1)): (68.0e6 + 31.0e6, "profiles/r02_ncu_conv2_fpn_output2_summary.txt")}
    traffic, traffic_src = ncu.get((kind, tuple(key)), (None, None))
    alg_bytes = 2.0 * (N * Hh * Ww * Cin + N * (Hh // stride) * (Ww // stride) * Cout + R * R * Cin * Cout)
    ach = flop / (ms * 1e-3) / 1e12
    name = {"fwd": "conv2_kernel (tcgen05 cta_group::2 implicit GEMM, forward)",
            "dgrad": "conv2_kernel (tcgen05 cta_group::2, input gradient, filter read MN-major)",
            "wgrad": "conv_wgrad2_kernel (tcgen05 cta_group::2, weight gradient)"}[kind]
    return {"bound": "tensor", "kernel": "%s, %dx%d conv %d->%d stride %d on %dx%dx%d px, %d launches per step"
                                         % (name, R, R, Cin, Cout, stride, N, Hh, Ww, cnt),
            "in_step": True, "achieved": ach, "peak": peaks["tf_sus"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sus"],
            "peak_source": peaks["src"] + " bf16 sustained (kernel timed inside the step)",
            "traffic": traffic, "traffic_unit": "bytes/launch", "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": alg_bytes,
            "algorithmic_flops_per_launch": flop / cnt, "ms_per_launch": ms / cnt,
            "share_of_step_time": ms / ms_step,
            "all_tcgen05_launches": {"launches_per_step": len(recs), "flop_per_step": tot_flop, "ms_per_step": tot_ms,
                                     "achieved": tot_flop / (tot_ms * 1e-3) / 1e12, "unit": "TFLOP/s",
                                     "frac": tot_flop / (tot_ms * 1e-3) / 1e12 / peaks["tf_sus"],
                                     "share_of_step_time": tot_ms / ms_step,
                                     "share_of_step_flop": tot_flop / (FLOP_PER_IMAGE * IMS_PER_GPU)},
            "groups": [{"kind": k[0], "shape_N_H_W_Cin_Cout_k_stride": list(k[1]), "launches": v[0], "ms": round(v[1], 4),
                        "tflops": round(v[2] / (v[1] * 1e-3) / 1e12, 1)}
                       for k, v in sorted(groups.items(), key=lambda kv: -kv[1][1])[:40]],
            "method": "CUDA events on the launching stream around each tcgen05 launch of a single-stream execution of the "
                      "static step (%s): the production graph overlaps independent branches on side streams, which would "
                      "stretch per-kernel durations" % ("captured as a second CUDA graph with external event-record nodes, "
                      "replayed" if how == "graph" else how)}


def conv_tc_roofline(peaks):
    """The same kernel in isolation (micro-benchmark): FPN-output / RPN-head 3x3 (256->256 on 2x256x256 px, 154.6 GFLOP per
    launch = 2*M*Cout*9*Cin), back-to-back launches on distinct buffers larger than L2 in aggregate, CUDA events on the
    launch stream."""
    from .modeling import conv_tc as _ct
    conv2d_nhwc = (lambda x, w, s, p: _ct.conv2_nhwc(x, w, s, p)) if _ct.USE_CONV2 else _ct.conv2d_nhwc
    N, C, Hh, Ww = 2, 256, 256, 256
    xs = [torch.randn(N, C, Hh, Ww, device="cuda").bfloat16().contiguous(memory_format=torch.channels_last) for _ in range(4)]
    w = (torch.randn(C, 3, 3, C, device="cuda") * 0.02).bfloat16()
    for i in range(4):
        conv2d_nhwc(xs[i], w, 1, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 40
    e0.record()
    for i in range(reps):
        conv2d_nhwc(xs[i % 4], w, 1, 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    flops = 2.0 * N * Hh * Ww * C * C * 9
    ach = flops / (ms * 1e-3) / 1e12
    return {"bound": "tensor", "kernel": "%s<256,bf16> (tcgen05 implicit GEMM), FPN output2 / RPN p2 shape, isolated"
                                         % ("conv2_kernel" if _ct.USE_CONV2 else "conv_tc_kernel"),
            "in_step": False, "achieved": ach, "peak": peaks["tf_burst"], "unit": "TFLOP/s", "frac": ach / peaks["tf_burst"],
            "peak_source": peaks["src"] + " bf16 burst (kernel timed alone)",
            "traffic": None, "traffic_unit": "bytes/launch",
            "algorithmic_flops_per_launch": flops, "ms_per_launch": ms}


def cpu_threads():
    """torch CPU threads for the reference arm: all cores up to 32 (beyond that the reference's many small ATen ops
    get slower, not faster: measured 125 s/step at 128 threads vs ~12 s at 8 on the same code)."""
    return max(1, min(os.cpu_count() or 1, 32))


def cpu_train_sample(steps, n_images=1):
    """The oracle port of the reference step (fp32, torch CPU ops, all host threads) on a bounded sample:
    `n_images` 1024x1024 images per step, forward + backward."""
    import torch
    from oracle import detector_oracle as do
    torch.set_num_threads(cpu_threads())
    cfg = do.DetCfg(NUM_CLASSES, SEM_CLASSES)
    params = {k: (v.clone().requires_grad_(v.is_floating_point() and "running" not in k)) for k, v in
              do.init_params(cfg, 0).items()}
    data = do.synthetic_batch(n_images, H, W, NUM_CLASSES, SEM_CLASSES, seed=1234)
    ts = []
    for i in range(steps + 1):
        t0 = time.perf_counter()
        torch.manual_seed(i)
        loss = sum(do.forward_train(params, cfg, *data).values())
        loss.backward()
        for p in params.values():
            p.grad = None
        if i > 0:
            ts.append(time.perf_counter() - t0)
    t = sum(ts) / len(ts)
    return {"value": n_images / t, "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d x 1024x1024 image(s) per step, fwd+bwd of the oracle port of PanopticFPN (fp32), %d timed step(s)"
                      % (n_images, steps), "s_per_step": t}


def reference_line(args):
    """bench.py --impl reference: the reference's CPU path (oracle port: torch CPU ops = the library calls the reference
    itself makes; /root/reference does not exist on the GPU box) on the arm's own config: 2 images of 1024x1024 per step,
    forward + backward. Bounded: 1 untimed + at most 3 timed steps (~10 s each on 32 threads); `steps` is what was timed."""
    timed = max(1, min(args.steps, 3))
    r = cpu_train_sample(timed, n_images=IMS_PER_GPU)
    return {"impl": "reference", "metric": "u2seg_R50_800_train_images_per_sec", "value": r["value"], "unit": "images/s",
            "n_gpus": args.gpus, "steps": timed, "warmup": 1, "ms_per_step": r["s_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "u2seg_R50_800.yaml training step (PanopticFPN R50-FPN, SyncBN, CascadeROIHeads, "
                                   "800 classes), batch 2/GPU, synthetic 1024x1024 COCO-panoptic-shaped inputs, fwd+bwd "
                                   "(oracle port of the reference's CPU path, fp32)", "global_batch": IMS_PER_GPU,
                       "image_size": [H, W], "requested_steps": args.steps, "requested_warmup": args.warmup},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
