"""Feature-space kNN timing on real layer activations (x1, x2 of the fused encoder on synthetic clouds),
pruned vs full tensor-core scan.  GCANET_TC_STATS=1 prints the tiles visited."""
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
feats = {'x1': x1.contiguous(), 'x2': x2.contiguous(), 'randn': torch.randn(B, 64, N, device='cuda')}
for name, f in feats.items():
    for prune in (True, False):
        for _ in range(3):
            G.knn_graph(f, k, k, want64=False, want32=True, ordered=False, prune=prune)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            G.knn_graph(f, k, k, want64=False, want32=True, ordered=False, prune=prune)
        b.record(); torch.cuda.synchronize()
        print(f'{name} prune={prune}: {a.elapsed_time(b) / 10:.3f} ms per call', flush=True)
