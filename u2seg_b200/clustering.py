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
    x16 = x.to(device=device, dtype=torch.float16, non_blocking=True).contiguous()
    # The reference keeps fp32 (nn_utils.py:304); here distances AND centroid means are taken over the fp16-rounded rows
    # (relative rounding 2^-11 per component; DINO embeddings are L2-normalised). Values beyond fp16's range would turn
    # into inf silently - refuse them instead.
    if x.dtype != torch.float16 and not bool(torch.isfinite(x16).all()):
        raise ValueError("u2seg_b200.clustering: embeddings exceed the fp16 range (|x| > 65504) or are not finite; "
                         "rescale them (k-means labels are invariant to a common scale)")
    return x16


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


def run_kMeans(feats_list, num_centroids, final_sample_num, train_memory_dataset=None, Niter=100, recompute=False,
               use_cuda=True, seed=None, force_no_lazy_tensor=False, save=True, save_dir=None):
    """nn_utils.py:382-405 `run_kMeans`, same arguments and defaults. recompute=True: k-means, then (save=True) the labels /
    centroids go to cluster_labels_{n}{_seed}.npy / centroids_{n}{_seed}.npy; recompute=False (the reference's default):
    those files are loaded. Either way cluster_labels_decode.json ({image path: cluster id}) is written when
    `train_memory_dataset` (an object with `.imgs`, as in the reference) is given. Existing files are never overwritten
    (nn_utils.py:75-107). `save_dir` stands in for the reference's global cfg.RUN_DIR; without it nothing touches the disk
    (= cfg.SKIP_SAVE) and recompute=False is an error. use_cuda / force_no_lazy_tensor: accepted, there is one device path."""
    import json
    import os

    import numpy as np
    sfx = "_{}".format(seed) if seed is not None else ""
    names = ("cluster_labels_{}{}.npy".format(final_sample_num, sfx), "centroids_{}{}.npy".format(final_sample_num, sfx))

    def save_once(name, writer):
        path = os.path.join(save_dir, name)
        if os.path.exists(path):
            print("File exists: {}. Not overwriting (if the file is stale, please save manually).".format(path))
        else:
            writer(path)
            print("File saved to: {}".format(path))

    if recompute:
        cluster_labels, centroids = KMeans(feats_list, seed=seed, K=num_centroids, Niter=Niter, verbose=True,
                                           force_no_lazy_tensor=force_no_lazy_tensor)
        cluster_labels, centroids = cluster_labels.cpu(), centroids.cpu()
        inds, cnts = torch.unique(cluster_labels, return_counts=True)
        print("Num of clusters: {} min: {} max: {}".format(len(inds), cnts.min().item(), cnts.max().item()))
        if save and save_dir is not None:
            save_once(names[0], lambda path: np.save(path, cluster_labels.numpy()))
            save_once(names[1], lambda path: np.save(path, centroids.numpy()))
    else:
        if save_dir is None:
            raise ValueError("run_kMeans(recompute=False) loads {} / {} from save_dir (the reference's cfg.RUN_DIR): "
                             "pass save_dir, or recompute=True".format(*names))
        cluster_labels = torch.tensor(np.load(os.path.join(save_dir, names[0])))
        centroids = torch.tensor(np.load(os.path.join(save_dir, names[1])))
    if train_memory_dataset is not None and save_dir is not None:
        def image_key(entry):            # nn_utils.py:87-91: the last two path components of the image file
            parts = entry[0].split("/")
            return "/".join(parts[-2:])

        decode = {image_key(train_memory_dataset.imgs[i]): int(cid) for i, cid in enumerate(cluster_labels.tolist())}
        save_once("cluster_labels_decode.json", lambda path: json.dump(decode, open(path, "w")))
    return cluster_labels, centroids


# ---------------------------------------------------------------------------------------------------------------------
# k-nearest neighbours (nn_utils.py:203-299): exact, tensor-core candidate pass + fp32 certification (csrc/knn.cu)
# ---------------------------------------------------------------------------------------------------------------------
def _knn_prepare(x32):
    """fp32 (N, D) CUDA rows -> (fp16 rows padded to the train-tile multiple, |x|^2 fp32 (+inf in the padding), max |x|^2).
    Use _knn_rounding(x32, x16) for the rows' fp16 rounding-error norm."""
    import ctypes
    L = _lib.lib()
    N, D = x32.shape
    npad = int(L.u2b_knn_npad(N))
    x16 = torch.empty((npad, D), dtype=torch.float16, device=x32.device)
    xn = torch.empty((npad,), dtype=torch.float32, device=x32.device)
    xmax2 = torch.zeros((1,), dtype=torch.float32, device=x32.device)
    # u2b_kmeans_prepare pads to ITS tile multiple (same 160-row tile as the kNN kernel): kpad == npad
    assert int(L.u2b_kmeans_kpad(N)) == npad
    _lib.check(L.u2b_kmeans_prepare(_lib.ptr(x32), N, D, _lib.ptr(x16), _lib.ptr(xn), _lib.ptr(xmax2), _lib.stream_ptr()),
               "u2b_kmeans_prepare")
    _lib.count_launches(1)
    return x16, xn, xmax2


def _knn_rounding(x32, x16, chunk=262144):
    """max_i |x_i - fp16(x_i)|_2 (the MEASURED rounding error of the candidate pass's operands), as a 0-d tensor."""
    worst = torch.zeros((), dtype=torch.float32, device=x32.device)
    n = x32.shape[0]                  # x16 may carry zero padding rows beyond n
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        worst = torch.maximum(worst, (x32[s:e] - x16[s:e].float()).norm(dim=1).max())
    return worst


def _knn_eps(xmax2, ymax2, ex, ey, D):
    """Bound on |candidate-pass value - (|y|^2 - 2 x.y)| for every pair: with x~, y~ the fp16 operands,
    |x.y - x~.y~| <= |x| |y - y~| + |x - x~| |y~|  (Cauchy-Schwarz; the rounding-error norms are measured, not the 2^-11
    worst case), plus the fp32 accumulation of D products and of |y|^2."""
    xm, ym = float(torch.sqrt(xmax2)), float(torch.sqrt(ymax2))
    ex, ey = float(ex), float(ey)
    return 2.0 * (xm * ey + ex * (ym + ey)) + 4.0 * D * 2.0 ** -24 * max(xm, ym) ** 2


def _knn_second_pass(xq, xnq, y, yn, K, M=128, rows_per_chunk=1024):
    """Rows the 48-candidate lists could not certify (dense neighbourhoods: more than ~48 train rows within the fp16
    rounding bound of the K-th distance). Values |y|^2 - 2 x.y for ALL train rows from an fp32 library GEMM (no fp16
    rounding, error <= eps2 = 4 D 2^-24 |x| |y|), the M smallest re-evaluated exactly, the same certificate with eps2.
    Returns (d (r,K), i (r,K), certified (r,) bool)."""
    N2, D = y.shape
    eps2 = 4.0 * D * 2.0 ** -24 * float(torch.sqrt(xnq.max() * yn[:N2].max()))
    d_out = torch.empty((xq.shape[0], K), dtype=torch.float32, device=xq.device)
    i_out = torch.empty((xq.shape[0], K), dtype=torch.int64, device=xq.device)
    ok = torch.empty((xq.shape[0],), dtype=torch.bool, device=xq.device)
    M = min(M, N2)
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        for s in range(0, xq.shape[0], rows_per_chunk):
            q = xq[s:s + rows_per_chunk]
            v = torch.addmm(yn[None, :N2].expand(q.shape[0], -1), q, y.t(), alpha=-2.0)
            vals, idx = torch.topk(v, M, dim=1, largest=False, sorted=True)
            del v
            idx, order = torch.sort(idx, dim=1)                       # by train index, so that the stable sort breaks ties by index
            d = ((q[:, None, :] - y[idx]) ** 2).sum(-1)
            d, o2 = torch.sort(d, dim=1, stable=True)
            d_out[s:s + q.shape[0]] = d[:, :K]
            i_out[s:s + q.shape[0]] = torch.gather(idx, 1, o2[:, :K])
            # every discarded row has value >= vals[:, M-1], i.e. true distance >= vals[:, M-1] + |x|^2 - eps2
            ok[s:s + q.shape[0]] = (d[:, K - 1] + eps2 < vals[:, M - 1] + xnq[s:s + q.shape[0]] - eps2) if M < N2 else True
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    return d_out, i_out, ok


def _knn_exhaustive(xq, y, K, chunk=65536):
    """exact fp32 kNN of a FEW query rows by the reference's dense formula (rows the fast path could not certify)."""
    best_d = torch.full((xq.shape[0], K), float("inf"), device=xq.device)
    best_i = torch.zeros((xq.shape[0], K), dtype=torch.int64, device=xq.device)
    for s in range(0, y.shape[0], chunk):
        d = ((xq[:, None, :] - y[None, s:s + chunk, :]) ** 2).sum(-1)
        dd = torch.cat([best_d, d], 1)
        ii = torch.cat([best_i, torch.arange(s, s + d.shape[1], device=xq.device).expand(xq.shape[0], -1)], 1)
        # stable sort by (distance, index): indices ascend within each block and blocks arrive in order
        vals, order = torch.sort(dd, dim=1, stable=True)
        best_d, best_i = vals[:, :K], torch.gather(ii, 1, order[:, :K])
    return best_d, best_i


def kNN(x_train, x_test, K=20, query_chunk=262144, return_stats=False):
    """nn_utils.py:203-224: (ind_knn int64 (N_test, K), d_knn fp32 (N_test, K)) = the K smallest squared L2 distances of
    every test row to the train rows, ascending, with the train indices. Exact fp32 results (sum_d (x-y)^2 evaluated in
    fp32 for the returned pairs); equal distances are ordered by index. x_*: (N, D) float tensors, D in {128, 256, 384}."""
    import ctypes
    L = _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    y = x_train.to(dev, torch.float32).contiguous()
    xt = x_test.to(dev, torch.float32).contiguous()
    assert y.dim() == 2 and xt.dim() == 2 and y.shape[1] == xt.shape[1]
    N2, D = y.shape
    assert N2 >= K, "kNN: fewer train rows than K"
    y16, yn, ymax2 = _knn_prepare(y)
    ey = _knn_rounding(y, y16)
    same = x_train is x_test or (xt.data_ptr() == y.data_ptr() and xt.shape == y.shape)
    NC = int(L.u2b_knn_candidates_per_row())
    ind = torch.empty((xt.shape[0], K), dtype=torch.int64, device=dev)
    dist = torch.empty((xt.shape[0], K), dtype=torch.float32, device=dev)
    n_flag_total = n_exhaustive = 0
    for s in range(0, xt.shape[0], query_chunk):
        xq = xt[s:s + query_chunk]
        n1 = xq.shape[0]
        if same:
            x16, xn, xmax2, ex = y16[s:s + n1], yn[s:s + n1], ymax2, ey
        else:
            x16, xn, xmax2 = _knn_prepare(xq)
            ex = _knn_rounding(xq, x16)
        cand = torch.empty((n1, NC), dtype=torch.int32, device=dev)
        cval = torch.empty((n1, NC), dtype=torch.float32, device=dev)
        thr = torch.empty((n1, 2), dtype=torch.float32, device=dev)
        _lib.check(L.u2b_knn_candidates(_lib.ptr(x16), n1, _lib.ptr(y16), _lib.ptr(yn), N2, D, _lib.ptr(cand), _lib.ptr(cval),
                                        _lib.ptr(thr), _lib.stream_ptr()), "u2b_knn_candidates")
        eps = _knn_eps(xmax2, ymax2, ex, ey, D)
        flagged = torch.empty((n1,), dtype=torch.int32, device=dev)
        nflag = torch.zeros((1,), dtype=torch.int32, device=dev)
        _lib.check(L.u2b_knn_refine(_lib.ptr(xq), _lib.ptr(y), _lib.ptr(cand), _lib.ptr(cval), _lib.ptr(thr), _lib.ptr(xn), n1, D, int(K),
                                    ctypes.c_float(eps), _lib.ptr(dist[s:s + n1]), _lib.ptr(ind[s:s + n1]), _lib.ptr(flagged),
                                    _lib.ptr(nflag), _lib.stream_ptr()), "u2b_knn_refine")
        _lib.count_launches(2)
        nf = int(nflag)
        if nf:        # not certifiable from 48 candidates (dense neighbourhood of the K-th neighbour)
            rows = flagged[:nf].long()
            d2, i2, ok = _knn_second_pass(xq[rows], xn[rows], y, yn, K)
            dist[s + rows] = d2
            ind[s + rows] = i2
            bad = rows[~ok]
            for b in range(0, bad.numel(), 512):     # exact ties beyond the second pass's 128 candidates: exhaustive
                rr = bad[b:b + 512]
                d_e, i_e = _knn_exhaustive(xq[rr], y, K)
                dist[s + rr] = d_e
                ind[s + rr] = i_e
            n_flag_total += nf
            n_exhaustive += int(bad.numel())
    if return_stats:
        return ind, dist, {"uncertified_rows": n_flag_total, "exhaustive_rows": n_exhaustive, "eps": eps}
    return ind, dist


def partitioned_kNN(feats_list, K=20, recompute=True, partitions_size=130000, verify=False, save_dir=None):
    """nn_utils.py:227-299: self-kNN of `feats_list` -> (d_knns (N,K) fp32, ind_knns (N,K) int64). The reference
    partitions train and test rows to bound KeOps' memory and merges the per-partition lists; the kernel here streams
    the whole train set past every query tile, so partitions_size only bounds the per-launch query chunk. With
    recompute=False the saved .npy files are loaded, as in the reference."""
    import os

    import numpy as np
    suffix = "" if K == 20 else "_{}".format(K)
    if not recompute:
        assert save_dir is not None, "recompute=False needs save_dir (where d_knns / ind_knns .npy live)"
        return (torch.tensor(np.load(os.path.join(save_dir, "d_knns{}.npy".format(suffix)))),
                torch.tensor(np.load(os.path.join(save_dir, "ind_knns{}.npy".format(suffix)))))
    ind, d = kNN(feats_list, feats_list, K=K, query_chunk=max(int(partitions_size), 1024))
    if verify:        # the reference's own check: distances of the selected neighbours are the true ones
        g = torch.Generator().manual_seed(0)
        rows = torch.randperm(feats_list.shape[0], generator=g)[:256].to(ind.device)
        d_e, _ = _knn_exhaustive(feats_list.to(ind.device, torch.float32)[rows], feats_list.to(ind.device, torch.float32), K)
        assert torch.allclose(d[rows], d_e, rtol=1e-5, atol=1e-6)
    d, ind = d.cpu(), ind.cpu()
    if save_dir is not None:
        np.save(os.path.join(save_dir, "d_knns{}.npy".format(suffix)), d.numpy())
        np.save(os.path.join(save_dir, "ind_knns{}.npy".format(suffix)), ind.numpy())
    return d, ind
