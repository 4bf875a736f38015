"""Small shapes through every kernel variant of the feature-space kNN and EdgeConv paths in one process (a quick
exerciser; compute-sanitizer is not available on the GPU pool)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import gcanet_b200 as gb
from gcanet_b200 import functional as G
from gcanet_b200.synth import abc_like_batch
torch.manual_seed(0)
B, N, k = 2, 2048, 20
enc = gb.DGCNNEncoderGn(mode=0, nn_nb=k, input_channels=6).cuda()
x = torch.from_numpy(abc_like_batch(B, N, seed=3)).cuda()
cot = [torch.randn(B, c, N, device='cuda') for c in (64, 64, 128)]
outs = enc.edge_stack(x)                       # xyz kNN, pruned feature kNN (SM=1), EdgeConv fwd incl. gemm_tc
torch.autograd.backward(outs, cot)             # EdgeConv bwd incl. small variant, gemm_tc dX, wide tn GEMM
x1 = outs[0].detach().contiguous()
G.knn_graph(x1, 80, 80)                        # SM=2 variant, ordered
y = x1.clone(); y[:, :, :300] = y[:, :, :1]
G.knn_graph(y, k, k)                           # overflow rows -> per-row fallback
G.knn_graph(x1, k, k, prune=False)             # full scan
torch.cuda.synchronize()
print('done')
