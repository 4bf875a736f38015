import sys, torch, time
sys.path.insert(0, '/root/repo')
from gcanet_b200 import functional as G
torch.manual_seed(0)
x = torch.randn(16, 64, 10000, device='cuda')
for _ in range(3): G.knn_graph(x, 50, 50, want64=False, want32=True, ordered=False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): G.knn_graph(x, 50, 50, want64=False, want32=True, ordered=False)
b.record(); torch.cuda.synchronize()
print('knn_graph C=64 total ms per call', a.elapsed_time(b) / 10)
