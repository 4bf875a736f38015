"""fwd+bwd of the three-layer kNN + EdgeConv stack for a given k / mode (CUDA events), e.g. the reference's
default k = 80:  python tools/time_stack.py 80 0"""
import sys, torch
sys.path.insert(0, '/root/repo')
import gcanet_b200 as gb
from gcanet_b200.synth import abc_like_batch
k = int(sys.argv[1]) if len(sys.argv) > 1 else 80
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B, N = 16, 10000
torch.manual_seed(0)
enc = gb.DGCNNEncoderGn(mode=mode, nn_nb=k, input_channels=6).cuda()
x = torch.from_numpy(abc_like_batch(B, N, seed=1234, with_normals=(mode == 5))).cuda()
cot = [torch.randn(B, c, N, device='cuda') for c in (64, 64, 128)]
def step():
    for p in enc.parameters(): p.grad = None
    outs = enc.edge_stack(x)
    torch.autograd.backward(outs, cot)
for _ in range(3): step()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): step()
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 10
print(f'k={k} mode={mode}: {ms:.3f} ms per step, {B / ms * 1e3:.0f} clouds/s')
