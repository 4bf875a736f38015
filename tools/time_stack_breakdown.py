import sys, torch
sys.path.insert(0, '/root/repo')
import gcanet_b200 as gb
from gcanet_b200 import functional as G
from gcanet_b200.synth import abc_like_batch
k = int(sys.argv[1]) if len(sys.argv) > 1 else 50
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 5
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
G.enable_kernel_timing(True)
for _ in range(5): step()
torch.cuda.synchronize()
for tag, v in sorted(G.kernel_timings_ms().items()): print(f'{tag}: {sum(v)/5:.3f} ms per step')
