"""Debug aid: EdgeConv backward with and without the tensor-core GEMMs on one shape; prints where they differ."""
import os, sys, torch
sys.path.insert(0, '/root/repo')
import gcanet_b200 as gb
from gcanet_b200 import functional as G
C, Cout, N, k, B = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 128, int(sys.argv[2]) if len(sys.argv) > 2 else 4501, 8, 2
g = torch.Generator().manual_seed(1)
x = torch.randn(B, C, N, generator=g).cuda()
W = (torch.randn(Cout, 2 * C, generator=g) / (2 * C) ** 0.5).cuda()
gamma = (torch.randn(Cout, generator=g) * 0.7 + 0.2).cuda()
beta = (torch.randn(Cout, generator=g) * 0.3).cuda()
cot = torch.randn(B, Cout, N, generator=g).cuda()
_, idx = G.knn_graph(x, k, k, want64=False, want32=True)
def run():
    xg, Wg, gg, bg = (t.clone().requires_grad_(True) for t in (x, W, gamma, beta))
    x_nc = G._ToPointMajor.apply(xg, C)
    out_nc, out_cn = gb.edgeconv(x_nc, idx, Wg, gg, bg, C, groups=2)
    (out_cn * cot).sum().backward()
    torch.cuda.synchronize()
    return out_cn.detach(), xg.grad, Wg.grad, gg.grad, bg.grad
a = run()
os.environ['GCANET_NO_TC_GEMM'] = '1'
b = run()
for name, u, v in zip(("out", "dx", "dW", "dgamma", "dbeta"), a, b):
    d = (u - v).abs()
    print(name, "max abs diff", float(d.max()), "ref max", float(v.abs().max()))
    if name == "dx":
        per_pt = d.amax(dim=1)    # [B, N]
        bad = (per_pt > 1e-3 * float(v.abs().max())).nonzero()
        print("  bad points:", bad.shape[0], bad[:10].tolist(), bad[-5:].tolist())
    if name == "dW":
        bad = (d > 1e-3 * float(v.abs().max())).nonzero()
        print("  bad entries:", bad.shape[0], bad[:10].tolist())
