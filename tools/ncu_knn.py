"""Developer tool for `ncu --set full -k regex:knn_candidates_kernel`: one wave (148 query tiles) of the kNN candidate pass
against 1.28 M train rows, D = 384. usage: python tools/ncu_knn.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from u2seg_b200 import _lib
from u2seg_b200.clustering import _knn_prepare
N, D, n1 = 1280000, 384, 148 * 128
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.nn.functional.normalize(torch.randn(N, D, generator=g, device="cuda"), dim=1)
y16, yn, _ = _knn_prepare(x)
L = _lib.lib()
nc = int(L.u2b_knn_candidates_per_row())
cand = torch.empty((n1, nc), dtype=torch.int32, device="cuda"); cval = torch.empty((n1, nc), device="cuda"); thr = torch.empty((n1, 2), device="cuda")
for _ in range(2):
    _lib.check(L.u2b_knn_candidates(_lib.ptr(y16), n1, _lib.ptr(y16), _lib.ptr(yn), N, D, _lib.ptr(cand), _lib.ptr(cval), _lib.ptr(thr), _lib.stream_ptr()), "cand")
torch.cuda.synchronize()
print("done")
