"""ORACLE (test infrastructure, never shipped): CPU restatement of the reference kNN.

Follows u2seg/Instance_Clustering/shared/utils/nn_utils.py:203-224:
  :209-213 D_ij = ((x_test[:, None, :] - x_train[None, :, :]) ** 2).sum(-1)      (fp32, squared L2)
  :216     d_knn, ind_knn = D_ij.Kmin_argKmin(K, dim=1)                          (K smallest per test row, ascending)
KeOps evaluates the reduction lazily; this restatement materialises D_ij in chunks of test rows and takes
torch.topk(largest=False, sorted=True). Pinned against the reference itself: oracle/make_golden.py runs the unmodified
nn_utils.kNN with a dense pykeops stub (oracle/ref_stubs/pykeops/torch.py) and tests/test_kmeans_oracle.py compares.
Ties: the reference leaves the order of equal distances to KeOps (its own verification, nn_utils.py:279-293, accepts any
index whose distance is right); comparisons therefore go through the distances and the index SETS.
"""
import torch


def knn_oracle(x_train, x_test, K=20, chunk=512):
    ind = torch.empty((x_test.shape[0], K), dtype=torch.int64)
    dist = torch.empty((x_test.shape[0], K), dtype=torch.float32)
    for s in range(0, x_test.shape[0], chunk):
        d = ((x_test[s:s + chunk, None, :] - x_train[None, :, :]) ** 2).sum(-1)
        v, i = torch.topk(d, K, dim=1, largest=False, sorted=True)
        dist[s:s + chunk], ind[s:s + chunk] = v, i
    return ind, dist
