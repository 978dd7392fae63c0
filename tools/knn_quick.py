"""Developer tool: candidate pass of the exact kNN (csrc/knn.cu) per cluster size, and the full search on a query chunk."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2seg_b200 import _lib
from u2seg_b200.clustering import _knn_prepare, kNN
N, D, n1 = int(sys.argv[1]) if len(sys.argv) > 1 else 1280000, 384, 131072
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.nn.functional.normalize(torch.randn(N, D, generator=g, device="cuda") + 2 * torch.randn(1000, D, generator=g, device="cuda")[torch.randint(0, 1000, (N,), generator=g, device="cuda")], dim=1)
y16, yn, _ = _knn_prepare(x)
L = _lib.lib()
nc = int(L.u2b_knn_candidates_per_row())
cand = torch.empty((n1, nc), dtype=torch.int32, device="cuda"); cval = torch.empty((n1, nc), device="cuda"); thr = torch.empty((n1, 2), device="cuda")
for cl in (1, 2, 4):
    _lib.check(L.u2b_knn_set_cluster(cl), "set_cluster")
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(L.u2b_knn_candidates(_lib.ptr(y16), n1, _lib.ptr(y16), _lib.ptr(yn), N, D, _lib.ptr(cand), _lib.ptr(cval), _lib.ptr(thr), _lib.stream_ptr()), "cand")
        e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("cluster %d: candidates %d x %d: %.1f ms = %.1f TF/s" % (cl, n1, N, ms, 2.0 * n1 * N * D / ms / 1e9), flush=True)
_lib.check(L.u2b_knn_set_cluster(int(os.environ.get("KNN_CL", "2"))), "set_cluster")
torch.cuda.synchronize(); t0 = time.perf_counter()
ind, d, st = kNN(x, x[:n1], K=20, return_stats=True)
torch.cuda.synchronize()
print("full kNN of %d queries: %.1f ms, stats %s" % (n1, (time.perf_counter() - t0) * 1e3, st))
