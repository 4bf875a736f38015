"""xyz kNN (C = 3, B = 16 x 10k, k = 50, set-only lists) a few times: for launch lists / ncu captures / the scan's counters
(GCANET_TC_STATS=1, GCANET_TC_PROF=1 in --aids builds).  Usage: python tools/xyz_tc_once.py [calls] [N] [B]"""
import sys
import torch
sys.path.insert(0, '/root/repo')
from gcanet_b200 import functional as G
from gcanet_b200.synth import abc_like_batch
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 3
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
B = int(sys.argv[3]) if len(sys.argv) > 3 else 16
x = torch.from_numpy(abc_like_batch(B, N, seed=1234)).cuda()
for _ in range(calls):
    G.knn_graph(x, 50, 50, G.METRIC_L2, want64=False, want32=True, ordered=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(calls):
    G.knn_graph(x, 50, 50, G.METRIC_L2, want64=False, want32=True, ordered=False)
b.record()
torch.cuda.synchronize()
print(f"xyz kNN B={B} N={N} k=50: {a.elapsed_time(b) / calls:.3f} ms per call")
