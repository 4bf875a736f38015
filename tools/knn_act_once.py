"""One pruned feature-space kNN call on layer-1 activations (for ncu)."""
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
x1 = x1.contiguous()
torch.cuda.synchronize()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    G.knn_graph(x1, k, k, want64=False, want32=True, ordered=False)
torch.cuda.synchronize()
