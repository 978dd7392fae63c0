"""k-means pseudo-label clustering — host-side mirror of
u2seg/Instance_Clustering/shared/utils/nn_utils.py:304-405 (`KMeans`, `run_kMeans`).

Same call signature and return values as the reference `KMeans`; the Lloyd iterations run in
libu2b200.so (tcgen05 distance GEMM + argmin epilogue, exact fp32 refinement, shared-memory
M-step). Embeddings are held in HBM as fp16 (N x D); centroids stay fp32. Row-sharded multi-GPU
(one process per GPU) adds one all-reduce of the (K, D+1) sums per iteration.
"""
import time

import torch

from . import _lib


def set_cluster(cl):
    """thread-block cluster size of the E-step kernel (1 = no multicast, 2, 4)."""
    _lib.check(_lib.lib().u2b_kmeans_set_cluster(int(cl)), "u2b_kmeans_set_cluster")


class KMeansState:
    """Device buffers for one (N, D, K) problem; reused across iterations."""

    def __init__(self, x16, K):
        L = _lib.lib()
        assert x16.is_cuda and x16.dtype == torch.float16 and x16.dim() == 2
        self.x16 = x16.contiguous()
        self.N, self.D = self.x16.shape
        self.K = int(K)
        dev = x16.device
        self.kpad = int(L.u2b_kmeans_kpad(self.K))
        self.c16 = torch.empty((self.kpad, self.D), dtype=torch.float16, device=dev)
        self.cnorm = torch.empty((self.kpad,), dtype=torch.float32, device=dev)
        self.scal = torch.zeros((4,), dtype=torch.float32, device=dev)  # [xmax, cmax2, -, -]
        self.labels = torch.empty((self.N,), dtype=torch.int32, device=dev)
        self.sums = torch.empty((self.K, self.D + 1), dtype=torch.float32, device=dev)
        self.amb_count = torch.zeros((1,), dtype=torch.int32, device=dev)
        self.ws_bytes = int(L.u2b_kmeans_workspace_bytes(self.N, self.D, self.K))
        self.ws = torch.empty((self.ws_bytes,), dtype=torch.uint8, device=dev)
        _lib.check(L.u2b_kmeans_xnorm_max(_lib.ptr(self.x16), self.N, self.D, _lib.ptr(self.scal),
                                          _lib.stream_ptr()), "u2b_kmeans_xnorm_max")
        _lib.count_launches(1)

    def set_xmax(self, xmax_tensor):
        self.scal[0:1].copy_(xmax_tensor)

    def assign(self, c32):
        L = _lib.lib()
        s = _lib.stream_ptr()
        cm2 = self.scal[1:2]
        _lib.check(L.u2b_kmeans_prepare(_lib.ptr(c32), self.K, self.D, _lib.ptr(self.c16),
                                        _lib.ptr(self.cnorm), _lib.ptr(cm2), s), "u2b_kmeans_prepare")
        _lib.check(L.u2b_kmeans_assign(_lib.ptr(self.x16), self.N, self.D, self.K, _lib.ptr(self.c16),
                                       _lib.ptr(c32), _lib.ptr(self.cnorm), _lib.ptr(self.scal),
                                       _lib.ptr(cm2), _lib.ptr(self.labels), _lib.ptr(self.amb_count),
                                       _lib.ptr(self.ws), self.ws_bytes, s), "u2b_kmeans_assign")
        _lib.count_launches(3)
        return self.labels

    def accumulate(self):
        L = _lib.lib()
        _lib.check(L.u2b_kmeans_accumulate(_lib.ptr(self.x16), _lib.ptr(self.labels), self.N, self.D,
                                           self.K, _lib.ptr(self.sums), _lib.ptr(self.ws),
                                           self.ws_bytes, _lib.stream_ptr()), "u2b_kmeans_accumulate")
        _lib.count_launches(2)
        return self.sums

    def finalize(self, c32):
        L = _lib.lib()
        _lib.check(L.u2b_kmeans_finalize(_lib.ptr(self.sums), self.K, self.D, _lib.ptr(c32),
                                         _lib.stream_ptr()), "u2b_kmeans_finalize")
        _lib.count_launches(1)
        return c32

    def lloyd_iteration(self, c32, group=None):
        """One E+M step, in place on c32 (nn_utils.py:348-364)."""
        self.assign(c32)
        self.accumulate()
        if group is not None:
            torch.distributed.all_reduce(self.sums, group=group)
        self.finalize(c32)


def to_device_fp16(x, device=None):
    """(N, D) embeddings -> contiguous fp16 on the GPU (the layout the kernels read)."""
    if not torch.cuda.is_available():
        raise RuntimeError("u2seg_b200.clustering needs a CUDA device (no CPU fallback)")
    device = device or torch.device("cuda", torch.cuda.current_device())
    return x.to(device=device, dtype=torch.float16, non_blocking=True).contiguous()


def init_centroids_sharded(x_local, r, row_offset, group=None):
    """c = x[r].clone() (nn_utils.py:338) when the rows of x are sharded over ranks: every rank contributes the
    rows it owns, one all-reduce(SUM) assembles the (K, D) fp32 centroids on all ranks."""
    N, D = x_local.shape
    r_dev = r.to(x_local.device)
    local = (r_dev >= row_offset) & (r_dev < row_offset + N)
    c = torch.zeros((r.shape[0], D), dtype=torch.float32, device=x_local.device)
    c[local] = x_local[(r_dev[local] - row_offset)].float()
    if group is not None:
        torch.distributed.all_reduce(c, group=group)
    return c


def lloyd_loop(state, c, Niter, group=None):
    """Niter x (E-step, local M-step sums, all-reduce of the (K, D+1) sums, centroid update). `state` provides
    assign(c) / accumulate() -> sums / finalize(c) (KMeansState on the GPU)."""
    for _ in range(Niter):
        state.assign(c)
        sums = state.accumulate()
        if group is not None:
            torch.distributed.all_reduce(sums, group=group)
        state.finalize(c)
    return c


def KMeans(x, seed, K=10, Niter=10, init_inds=None, verbose=True, force_no_lazy_tensor=False,
           group=None, row_offset=0, n_global=None):
    """Lloyd's algorithm, Euclidean metric. Mirrors nn_utils.py:304 `KMeans`.

    x: (N, D) tensor (any device / float dtype; converted to fp16 in HBM). Returns
    (cl int64 (N,), c fp32 (K, D)) like the reference. `force_no_lazy_tensor` is accepted for
    signature compatibility (there is one device path). Row-sharded use: every rank passes its
    shard, `row_offset`/`n_global` describe its position and `group` the process group.
    """
    start = time.time()
    x16 = to_device_fp16(x)
    N, D = x16.shape
    n_glob = int(n_global) if n_global is not None else N
    if seed is not None:
        torch.manual_seed(seed)
        torch.cuda.manual_seed(seed)
    if init_inds is None:
        r = torch.randperm(n_glob)[:K]            # CPU generator, as nn_utils.py:316
    else:
        assert K <= init_inds.shape[0]
        r = torch.randperm(init_inds.shape[0])[:K]
        r = init_inds[r]
    assert r.shape[0] == K, "{} != {}".format(r.shape[0], K)
    if verbose:
        print("Init indices {}".format(r.numpy()))

    st = KMeansState(x16, K)
    c = init_centroids_sharded(x16, r, row_offset, group)     # c = x[r].clone(); rows may live on other ranks
    if group is not None:
        torch.distributed.all_reduce(st.scal[0:1], op=torch.distributed.ReduceOp.MAX, group=group)
    lloyd_loop(st, c, Niter, group)

    cl = st.labels.long()
    if verbose:
        torch.cuda.synchronize()
        end = time.time()
        print(f"K-means for the Euclidean metric with {n_glob:,} points in dimension {D:,}, K = {K:,}:")
        print("Timing for {} iterations: {:.5f}s = {} x {:.5f}s\n".format(
            Niter, end - start, Niter, (end - start) / max(Niter, 1)))
    return cl, c


def run_kMeans(feats_list, num_centroids, final_sample_num=None, train_memory_dataset=None, Niter=100,
               recompute=True, use_cuda=True, seed=None, force_no_lazy_tensor=False, save=False,
               save_dir=None):
    """Mirror of nn_utils.py:382 `run_kMeans` (recompute branch; optional .npy dump)."""
    cluster_labels, centroids = KMeans(feats_list, seed=seed, K=num_centroids, Niter=Niter, verbose=False)
    cluster_labels, centroids = cluster_labels.cpu(), centroids.cpu()
    if save and save_dir is not None:
        import os

        import numpy as np
        sfx = "_{}".format(seed) if seed is not None else ""
        np.save(os.path.join(save_dir, "cluster_labels_{}{}.npy".format(final_sample_num, sfx)),
                cluster_labels.numpy())
        np.save(os.path.join(save_dir, "centroids_{}{}.npy".format(final_sample_num, sfx)), centroids.numpy())
    return cluster_labels, centroids
