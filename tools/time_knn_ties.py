import sys, torch
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
def t(f):
    for _ in range(2): G.knn_graph(f, k, k, want64=False, want32=True, ordered=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): G.knn_graph(f, k, k, want64=False, want32=True, ordered=False)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 5
base = x1.contiguous()
print('no ties', t(base))
for m in (10, 300, 3000):
    y = base.clone(); y[:, :, :m] = y[:, :, :1]
    print(f'{m} coincident points per cloud:', t(y))
