import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from prographmsa_b200 import distance
N = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
rng = np.random.default_rng(1)
A = rng.random((N, N)); D = (A + A.T) / 2; np.fill_diagonal(D, 0)
B = rng.random((N, N)) * 1e-3; V = (B + B.T) / 2; np.fill_diagonal(V, 0)
distance.bionj_joins(D[:50, :50].copy(), V[:50, :50].copy())
ts = []
for _ in range(4):
    t0 = time.perf_counter(); distance.bionj_joins(D, V); ts.append(time.perf_counter() - t0)
print("N=%d: " % N + " ".join("%.3f" % t for t in ts) + " s")
