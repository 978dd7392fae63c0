"""Quick GPU timing of the k-means kernels (developer tool; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.kmeans_oracle import make_mixture
from u2seg_b200.clustering import KMeansState, set_cluster

N, D, K = int(sys.argv[1]) if len(sys.argv) > 1 else 1280000, 384, 800
x16 = make_mixture(N, D, 1000, seed=0, spread=1.0).cuda()
c = x16[torch.randperm(N)[:K].cuda()].float().contiguous()
st = KMeansState(x16, K)
for _ in range(3):
    st.lloyd_iteration(c)
torch.cuda.synchronize()
def timeit(fn, n=10):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for cl in (1, 2, 4):
    set_cluster(cl)
    tcl = timeit(lambda: st.assign(c))
    print(f"cluster {cl}: assign {tcl:.3f} ms = {2*N*K*D/tcl/1e9:.1f} TFLOP/s")
if len(sys.argv) > 2:
    set_cluster(int(sys.argv[2]))
ta = timeit(lambda: st.assign(c))
tm = timeit(lambda: st.accumulate())
tf = timeit(lambda: st.finalize(c))
ti = timeit(lambda: st.lloyd_iteration(c))
print(f"N={N} assign {ta:.3f} ms  accumulate {tm:.3f} ms finalize {tf:.3f} ms iter {ti:.3f} ms  amb={int(st.amb_count)}")
print(f"assign tensor TFLOP/s = {2*N*K*D/ta/1e9:.1f}; emb/s = {N/ti*1e3:.3e}")
