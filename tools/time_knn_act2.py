"""Time of the pruned feature-space kNN call on layer-1 activations (B=16 x 10k, k=50), median of 15."""
import sys, statistics, torch
sys.path.insert(0, '/root/repo')
import gcanet_b200 as gb
from gcanet_b200 import functional as G
from gcanet_b200.synth import abc_like_batch
torch.manual_seed(0)
B, N, k = 16, 10000, 50
enc = gb.DGCNNEncoderGn(mode=0, nn_nb=k, input_channels=6).cuda()
x = torch.from_numpy(abc_like_batch(B, N, seed=1234)).cuda()
with torch.no_grad():
    x1, x2, x3 = enc.edge_stack(x)
for name, t in (("x1", x1.contiguous()), ("x2", x2.contiguous())):
    for _ in range(3): G.knn_graph(t, k, k, want64=False, want32=True, ordered=False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(15):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); G.knn_graph(t, k, k, want64=False, want32=True, ordered=False); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"{name}: {statistics.median(ts):.4f} ms")
