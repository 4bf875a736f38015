"""kNN sweep timings (BASELINE configs[2] and [4]): C = 3 / 64 / 128 at N = 10k, B = 16, k = 20 / 50, and the
100k-point stress shape; Gaussian features (no structure to prune) and encoder activations."""
import sys, torch
sys.path.insert(0, '/root/repo')
import gcanet_b200 as gb
from gcanet_b200 import functional as G
from gcanet_b200.synth import abc_like_batch

def t(f, k, metric=0, **kw):
    for _ in range(2): G.knn_graph(f, k, k, metric, want64=False, want32=True, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): G.knn_graph(f, k, k, metric, want64=False, want32=True, **kw)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 5

torch.manual_seed(0)
B, N = 16, 10000
xyz = torch.from_numpy(abc_like_batch(B, N, seed=1234)).cuda()
enc = gb.DGCNNEncoderGn(mode=0, nn_nb=50, input_channels=6).cuda()
with torch.no_grad():
    x1, x2, x3 = enc.edge_stack(xyz)
for k in (20, 50):
    print(f'k={k}: xyz C=3 {t(xyz, k):.3f} ms | x1 C=64 {t(x1.contiguous(), k):.3f} | x3 C=128 {t(x3.contiguous(), k):.3f} | '
          f'randn C=64 {t(torch.randn(B, 64, N, device="cuda"), k):.3f} | randn C=128 {t(torch.randn(B, 128, N, device="cuda"), k):.3f}', flush=True)
big = torch.from_numpy(abc_like_batch(4, 100000, seed=7)).cuda()
with torch.no_grad():
    b1, b2, b3 = enc.edge_stack(big)
print(f'B=4 x 100k, k=50: xyz {t(big, 50):.2f} ms | x1 C=64 {t(b1.contiguous(), 50):.2f} ms | randn C=64 {t(torch.randn(4, 64, 100000, device="cuda"), 50):.2f} ms', flush=True)
