"""Experiment: does visiting points in a spatially sorted order help the EdgeConv gather/scatter (L1 reuse)?
Sorts every synthetic cloud by a 30-bit xyz Morton code before running the stack and prints the per-op timings
next to the unsorted ones."""
import sys, torch
sys.path.insert(0, '/root/repo')
import gcanet_b200 as gb
from gcanet_b200 import functional as G
from gcanet_b200.synth import abc_like_batch
k = int(sys.argv[1]) if len(sys.argv) > 1 else 50
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
B, N = 16, 10000
torch.manual_seed(0)
enc = gb.DGCNNEncoderGn(mode=mode, nn_nb=k, input_channels=6).cuda()
x0 = torch.from_numpy(abc_like_batch(B, N, seed=1234, with_normals=(mode == 5))).cuda()

def morton_sort(x):
    p = x[:, :3]
    lo = p.amin(dim=2, keepdim=True); hi = p.amax(dim=2, keepdim=True)
    q = ((p - lo) / (hi - lo + 1e-9) * 1023).long().clamp(0, 1023)
    def spread(v):
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    code = spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)
    perm = code.argsort(dim=1)
    return torch.gather(x, 2, perm[:, None, :].expand_as(x)).contiguous()

cot = [torch.randn(B, c, N, device='cuda') for c in (64, 64, 128)]
for name, x in (("unsorted", x0), ("morton-sorted", morton_sort(x0))):
    def step():
        for p in enc.parameters(): p.grad = None
        outs = enc.edge_stack(x)
        torch.autograd.backward(outs, cot)
    G.enable_kernel_timing(False)
    for _ in range(3): step()
    torch.cuda.synchronize()
    G.enable_kernel_timing(True)
    G.kernel_timings_ms().clear() if hasattr(G.kernel_timings_ms(), 'clear') else None
    for _ in range(5): step()
    torch.cuda.synchronize()
    print(name)
    tot = 0
    for tag, v in sorted(G.kernel_timings_ms().items()):
        print(f'  {tag}: {sum(v[-5:])/5:.3f} ms per step'); tot += sum(v[-5:]) / 5
    print(f'  total {tot:.3f}')
