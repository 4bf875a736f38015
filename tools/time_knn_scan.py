"""Per-kernel CUDA-event style timing is not available through the C-ABI, so: total knn_graph time on layer
activations for a few settings (env GCANET_TC_PRE etc. are read per call)."""
import os, sys, torch
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
x1 = x1.contiguous(); x2 = x2.contiguous()
def t(f, **kw):
    for _ in range(3): G.knn_graph(f, k, k, want64=False, want32=True, ordered=False, **kw)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): G.knn_graph(f, k, k, want64=False, want32=True, ordered=False, **kw)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 10
for pre in sys.argv[1:] or ['8']:
    os.environ['GCANET_TC_PRE'] = pre
    os.environ['GCANET_TC_STATS'] = '1'
    G.knn_graph(x1, k, k, want64=False, want32=True, ordered=False)
    torch.cuda.synchronize()
    os.environ['GCANET_TC_STATS'] = '0'
    print(f'pre={pre}: x1 {t(x1):.3f} ms  x2 {t(x2):.3f} ms', flush=True)
