#!/usr/bin/env python
"""bench.py — driver contract.

    python bench.py --gpus N --steps K --warmup W [--workload train|kmeans] [--impl reference]

Workloads (BASELINE.json `metric`: "u2seg_R50_800 train images/sec ...; k-means embeddings/sec"):
  train   u2seg_R50_800 training step, batch 2 / GPU, synthetic 1024x1024 (configs[1]) — default
          once the detector step is available in this build, else kmeans
  kmeans  Lloyd iterations, N=1.28M, D=384, K=800 (configs[3]); rows sharded over ranks
  infer   u2seg_R50_300 panoptic inference, synthetic 800x1333, batch 1 (configs[4])
  knn     exact self-kNN (K=20) over the same 1.28M x 384 embeddings, explicit workload only

One JSON line on rank 0. `value` = device-resident throughput (CUDA events, max over ranks);
`e2e` = same metric through the public API with host buffers (H2D/D2H inside the timed region);
`roofline` = dominant kernel vs MEASURED_PEAKS.json; `cpu_baseline` = the oracle port timed on
the host cores on a bounded sample. `--impl reference` times only that CPU port.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

KM_N, KM_D, KM_K = 1_280_000, 384, 800


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sus=d["bf16_tflops_sustained"],
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sus=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi sampling DURING the timed region (recipe in B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_info():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# --------------------------------------------------------------------------------------
# CPU arm: the oracle port of nn_utils.KMeans (the reference needs pykeops + a GPU; its dense
# branch needs 1.5 TB at this size), chunked, on all host threads.
# --------------------------------------------------------------------------------------
def kmeans_cpu_sample(rows, steps, warmup):
    import torch
    from oracle.kmeans_oracle import assign_oracle, make_mixture, update_oracle
    torch.set_num_threads(max(1, min(os.cpu_count() or 1, 32)))
    x = make_mixture(rows, KM_D, 1000, seed=0, spread=1.0).float()
    g = torch.Generator().manual_seed(0)
    c = x[torch.randperm(rows, generator=g)[:KM_K]].clone()
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        cl = assign_oracle(x, c, chunk=64)
        c2, _ = update_oracle(x, cl, KM_K)
        c = torch.where(torch.isnan(c2), c, c2)
        t1 = time.perf_counter()
        if i >= warmup:
            ts.append(t1 - t0)
    t = sum(ts) / len(ts)
    return rows / t, t, torch.get_num_threads()


def run_reference(args):
    rank, world, _ = dist_info()
    if rank != 0:
        return
    if args.workload == "kmeans":
        rows = 32768
        v, t, cores = kmeans_cpu_sample(rows, max(1, min(args.steps, 5)), 1)
        line = {"impl": "reference", "metric": "kmeans_embeddings_per_sec", "value": v, "unit": "embeddings/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": "kmeans Lloyd iteration N=1.28M D=384 K=800 (oracle port of nn_utils.KMeans, "
                                       "%d-row sample per step)" % rows},
                "cpu_baseline": {"value": v, "unit": "embeddings/s", "cores": cores, "kind": "port",
                                 "sample": "%d rows x K=800 x D=384, one E+M step" % rows},
                "e2e": {"value": v, "unit": "embeddings/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    else:
        from u2seg_b200.bench_train import reference_line
        line = reference_line(args)
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# B200 arm, k-means
# --------------------------------------------------------------------------------------
def run_kmeans(args, emit=True):
    import torch
    import torch.distributed as dist
    from u2seg_b200 import _lib
    from u2seg_b200.clustering import KMeans, KMeansState

    rank, world, local = dist_info()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    peaks = load_peaks()

    # rows sharded: rank r owns [r*n_loc, (r+1)*n_loc)
    n_loc = KM_N // world
    g = torch.Generator(device=dev).manual_seed(1234)
    centres = torch.randn(1000, KM_D, generator=g, device=dev)
    g2 = torch.Generator(device=dev).manual_seed(99 + rank)
    x16 = torch.empty((n_loc, KM_D), dtype=torch.float16, device=dev)
    for s in range(0, n_loc, 160000):   # chunked to bound the fp32 temporaries
        e = min(n_loc, s + 160000)
        which = torch.randint(0, 1000, (e - s,), generator=g2, device=dev)
        blk = torch.randn(e - s, KM_D, generator=g2, device=dev) + 1.0 * centres[which]
        x16[s:e] = torch.nn.functional.normalize(blk, dim=1).half()
    st = KMeansState(x16, KM_K)
    c = x16[torch.randperm(n_loc, generator=g2, device=dev)[:KM_K]].float().contiguous()
    if world > 1:
        dist.broadcast(c, 0)
        dist.all_reduce(st.scal[0:1], op=dist.ReduceOp.MAX)

    def step():
        st.lloyd_iteration(c, group=group)

    for _ in range(max(3, args.warmup)):
        step()
    sampler = ClockSampler(local)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = _lib.launch_count
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = _lib.launch_count - launches0
    t = torch.tensor([ms_total], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t) / args.steps
    value = n_loc * world / (ms_step * 1e-3)

    # dominant kernel (E-step: tcgen05 assign + refinement), timed live on the launching stream
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ea.record()
    for _ in range(args.steps):
        st.assign(c)
    eb.record()
    torch.cuda.synchronize()
    ms_assign = ea.elapsed_time(eb) / args.steps
    flops = 2.0 * n_loc * KM_K * KM_D
    achieved = flops / (ms_assign * 1e-3) / 1e12
    roof = {"bound": "tensor", "kernel": "kmeans_assign_kernel (+prepare, refine)", "achieved": achieved,
            "peak": peaks["tf_sus"], "unit": "TFLOP/s", "frac": achieved / peaks["tf_sus"],
            "peak_source": peaks["src"] + " bf16 sustained",
            # dram__bytes_read.sum + dram__bytes_write.sum of kmeans_assign_kernel, one ncu --set full capture at the
            # BASELINE size (profiles/r01_ncu_kmeans_assign_cl2.txt): 983.9 MB + 8.4 MB = the fp16 embeddings once
            "traffic": (992.3e6 if (n_loc, KM_D, KM_K) == (1280000, 384, 800) else None), "traffic_unit": "bytes/launch",
            "algorithmic_flops_per_launch": flops, "ms_per_launch": ms_assign}

    line = {"metric": "kmeans_embeddings_per_sec", "value": value, "unit": "embeddings/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16 operands, f32 accumulate + f32 refinement",
            "data": "synthetic",
            "config": {"workload": "Instance_Clustering k-means Lloyd iteration (E+M), N=1.28M D=384 K=800, "
                                   "rows sharded over ranks", "N": KM_N, "D": KM_D, "K": KM_K,
                       "l2_note": "X shard (%.0f MB) exceeds the 126 MB L2, re-read from HBM every step"
                                  % (n_loc * KM_D * 2 / 1e6), "parallelism": "rows/%d" % world},
            "clocks": clocks, "gpu_launches": launches, "roofline": roof}

    if rank == 0 or world > 1:
        # e2e: public API KMeans() with HOST (pinned) buffers; H2D of X and D2H of labels+centroids timed
        niter = 20
        xh = x16.cpu().pin_memory()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        reps = 2
        for rep in range(reps + 1):     # one untimed repetition first: pinned-buffer first touch, lazy allocations
            if rep == 1:
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                t0 = time.perf_counter()
            cl, cc = KMeans(xh, 0, K=KM_K, Niter=niter, verbose=False, group=group,
                            row_offset=rank * n_loc, n_global=n_loc * world)
            cl_h, cc_h = cl.cpu(), cc.cpu()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        tt = torch.tensor([dt], device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt)
        line["e2e"] = {"value": n_loc * world * niter / dt, "unit": "embeddings/s",
                       "h2d_bytes_per_step": int(xh.numel() * 2), "d2h_bytes_per_step": int(n_loc * 8 + KM_K * KM_D * 4),
                       "what": "KMeans(x_host_pinned_fp16, seed, K=800, Niter=%d) incl. H2D of X, D2H of labels+centroids; "
                               "value = N*Niter/time" % niter}
    if rank == 0 and world == 1 and not os.environ.get("U2B_BENCH_SKIP_CPU"):
        rows = 16384
        v, tcpu, cores = kmeans_cpu_sample(rows, 2, 1)
        line["cpu_baseline"] = {"value": v, "unit": "embeddings/s", "cores": cores, "kind": "port",
                                "sample": "%d rows x K=800 x D=384, one E+M step (oracle port of nn_utils.KMeans)" % rows}
    del st, x16
    torch.cuda.empty_cache()
    if not emit:
        return line
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


# --------------------------------------------------------------------------------------
# B200 arm, kNN (nn_utils.py:203-299, the density-peak selection's neighbour search)
# --------------------------------------------------------------------------------------
def run_knn(args, emit=True):
    import torch
    from u2seg_b200 import _lib
    from u2seg_b200.clustering import _knn_prepare, kNN

    rank, world, local = dist_info()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = load_peaks()
    N, D, K = KM_N, KM_D, 20
    n_q = N // world                                       # query rows sharded over ranks; every rank holds the train set
    g = torch.Generator(device=dev).manual_seed(1234)
    centres = torch.randn(1000, D, generator=g, device=dev)
    x = torch.empty((N, D), dtype=torch.float32, device=dev)
    for s in range(0, N, 160000):
        e = min(N, s + 160000)
        which = torch.randint(0, 1000, (e - s,), generator=g, device=dev)
        x[s:e] = torch.nn.functional.normalize(torch.randn(e - s, D, generator=g, device=dev) + centres[which], dim=1)
    xq = x[rank * n_q:(rank + 1) * n_q]
    kNN(x[:20000], x[:4096], K=K)                          # lazy initialisation (function attributes, allocator)
    torch.cuda.synchronize()
    l0 = _lib.launch_count
    sampler = ClockSampler(local)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ind, d, stats = kNN(x, xq, K=K, return_stats=True)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count - l0
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
    # the candidate pass alone (dominant kernel), timed live on a 131,072-query chunk
    y16, yn, _ = _knn_prepare(x)
    nc = int(_lib.lib().u2b_knn_candidates_per_row())
    n1 = 131072
    cand = torch.empty((n1, nc), dtype=torch.int32, device=dev)
    cval = torch.empty((n1, nc), dtype=torch.float32, device=dev)
    thr = torch.empty((n1, 2), dtype=torch.float32, device=dev)
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record()
    _lib.check(_lib.lib().u2b_knn_candidates(_lib.ptr(y16), n1, _lib.ptr(y16), _lib.ptr(yn), N, D, _lib.ptr(cand), _lib.ptr(cval),
                                             _lib.ptr(thr), _lib.stream_ptr()), "u2b_knn_candidates")
    eb.record()
    torch.cuda.synchronize()
    ms_c = ea.elapsed_time(eb)
    flops = 2.0 * n1 * N * D
    ach = flops / (ms_c * 1e-3) / 1e12
    line = {"metric": "knn_queries_per_sec", "value": n_q * world / (ms * 1e-3), "unit": "queries/s", "n_gpus": world,
            "steps": 1, "warmup": 1, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16 candidate pass (f32 accumulate) + f32 exact pass", "data": "synthetic",
            "config": {"workload": "exact self-kNN, K=20, N=1.28M L2-normalised embeddings, D=384 (nn_utils.partitioned_kNN), "
                                   "query rows sharded over ranks", "N": N, "D": D, "K": K,
                       "uncertified_rows_recomputed": stats["uncertified_rows"],
                       "l2_note": "train set (983 MB fp16 + 1.97 GB fp32) exceeds the 126 MB L2"},
            "clocks": clocks, "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "knn_candidates_kernel (tcgen05, 131072 queries x 1.28M train rows)",
                         "achieved": ach, "peak": peaks["tf_sus"], "unit": "TFLOP/s", "frac": ach / peaks["tf_sus"],
                         "peak_source": peaks["src"] + " bf16 sustained", "traffic": None,
                         "algorithmic_flops_per_launch": flops, "ms_per_launch": ms_c},
            "e2e": {"value": n_q * world / (ms * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0, "what": "device-resident embeddings in, device-resident (ind, dist) out"}}
    del x, y16, cand
    torch.cuda.empty_cache()
    if not emit:
        return line
    if rank == 0:
        print(json.dumps(line), flush=True)


def run_dino(args, emit=True):
    """DINO ViT-S/8 feature extraction (u2seg_b200/dino.py) at 480x480 (3601 tokens), 8 images per step."""
    import torch
    from u2seg_b200 import _lib
    from u2seg_b200.dino import vit_small
    from u2seg_b200.modeling import conv_tc

    rank, world, local = dist_info()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    peaks = load_peaks()
    B, HW, P = 8, 480, 8
    torch.manual_seed(0)
    model = vit_small(patch_size=P, num_classes=0).to(dev).eval()
    g = torch.Generator().manual_seed(1234 + rank)
    host = [torch.randn(B, 3, HW, HW, generator=g).pin_memory() for _ in range(3)]
    pool = [h.to(dev) for h in host]
    N = (HW // P) ** 2 + 1
    D, depth, hid = 384, 12, 1536
    gemm_flop = 2.0 * B * N * (3 * P * P * D / N * (N - 1) + depth * (3 * D * D + D * D + 2 * D * hid))
    attn_flop = depth * 4.0 * B * N * N * D
    warm = max(3, args.warmup)
    with torch.no_grad():
        for i in range(warm):
            model(pool[i % 3])
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        sampler.start()
        l0 = _lib.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            feats = model(pool[i % 3])
        e1.record()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        launches = _lib.launch_count - l0
        ms = e0.elapsed_time(e1) / args.steps
        # end to end: pinned host images in (H2D every step), features read back every step
        t0 = time.perf_counter()
        for i in range(args.steps):
            host_feats = model(host[i % 3].to(dev, non_blocking=True)).cpu()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.steps
        # the tensor-core GEMMs alone: CUDA events around every conv2 launch of one forward
        conv_tc.TIMING = []
        model(pool[0])
        torch.cuda.synchronize()
        recs, conv_tc.TIMING = conv_tc.TIMING, None
    gemm_ms = sum(r[3].elapsed_time(r[4]) for r in recs)
    gemm_f = sum(r[2] for r in recs)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=dev)
        t = torch.tensor([ms, dt * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, dt = float(t[0]), float(t[1]) * 1e-3
    line = {"metric": "dino_vits8_480_images_per_sec", "value": B * world / (ms * 1e-3), "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 GEMM operands (fp32 accumulate, fp32 residual stream and LayerNorm)",
            "data": "synthetic",
            "config": {"workload": "DINO ViT-S/8 feature extraction (selective_labeling/dino.py ViTFeat.forward), 480x480 -> 3601 "
                                   "tokens, 8 images per step per GPU, random weights", "tokens": N, "batch_per_gpu": B,
                       "flop_per_step": gemm_flop + attn_flop, "l2_note": "inputs rotate over 3 batches (66 MB); activations "
                                                                           "of one step (~0.5 GB) exceed the 126 MB L2"},
            "clocks": clocks, "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "conv2_kernel as the ViT's Linear layers (%d launches per forward)" % len(recs),
                         "achieved": gemm_f / (gemm_ms * 1e-3) / 1e12 if gemm_ms else None, "peak": peaks["tf_sus"],
                         "unit": "TFLOP/s", "frac": gemm_f / (gemm_ms * 1e-3) / 1e12 / peaks["tf_sus"] if gemm_ms else None,
                         "peak_source": peaks["src"] + " bf16 sustained", "traffic": None,
                         "ms_per_forward_in_gemms": gemm_ms, "gemm_flop_per_forward": gemm_f,
                         "whole_forward_tflops": (gemm_flop + attn_flop) / (ms * 1e-3) / 1e12,
                         "note": "attention (scaled_dot_product_attention), LayerNorm and GELU are library calls in this "
                                 "version of the row; the GEMMs are 99 % of the non-attention flop"},
            "e2e": {"value": B * world / dt, "unit": "images/s", "h2d_bytes_per_step": B * 3 * HW * HW * 4,
                    "d2h_bytes_per_step": int(host_feats.numel() * 4)}}
    if rank == 0 and world == 1 and not os.environ.get("U2B_BENCH_SKIP_CPU"):
        from oracle import dino_oracle as vo
        from u2seg_b200.bench_train import cpu_threads
        cfg = vo.ViTCfg(patch_size=P, embed_dim=D, depth=depth, num_heads=6)
        sd = vo.init_params(cfg, 0)
        x1 = vo.synthetic_images(1, HW, HW, 1)
        torch.set_num_threads(cpu_threads())
        with torch.no_grad():
            vo.forward_features(sd, cfg, x1[:, :, :96, :96])
            t0 = time.perf_counter()
            vo.forward_features(sd, cfg, x1)
            tc = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": 1.0 / tc, "unit": "images/s", "cores": cpu_threads(), "kind": "port",
                                "sample": "1 image of 480x480 through the oracle port of the reference ViT-S/8 (fp32, torch CPU)"}
    if not emit:
        return line
    if rank == 0:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=["train", "kmeans", "infer", "knn", "dino"])
    args = ap.parse_args()
    if args.workload is None:
        args.workload = "train" if os.path.exists(os.path.join(ROOT, "u2seg_b200", "bench_train.py")) else "kmeans"
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "kmeans":
        return run_kmeans(args)
    if args.workload == "knn":
        return run_knn(args)
    if args.workload == "dino":
        return run_dino(args)
    if args.workload == "infer":
        from u2seg_b200.bench_infer import run_infer
        return run_infer(args, ClockSampler, load_peaks, dist_info)
    from u2seg_b200.bench_train import run_train
    return run_train(args, ClockSampler, load_peaks, dist_info, run_kmeans)


if __name__ == "__main__":
    main()
