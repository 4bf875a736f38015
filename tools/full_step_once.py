"""A few per-point GCANet steps (stack + tail + heads + normals EdgeConv + offset block + losses) for ncu."""
import sys, torch
sys.path.insert(0, '/root/repo')
from gcanet_b200.model import PrimitivesEmbeddingPerPoint, nll_loss, offset_l1_loss
from gcanet_b200.synth import abc_like_batch
torch.manual_seed(0)
B, N = 16, 10000
net = PrimitivesEmbeddingPerPoint(mode=5, nn_nb=50).cuda()
c = torch.from_numpy(abc_like_batch(B, N, seed=4321, with_normals=True)).cuda()
pts, nrm = c[:, :3].transpose(1, 2).contiguous(), c[:, 3:].transpose(1, 2).contiguous()
g = torch.Generator().manual_seed(3)
t_gt = torch.randint(0, 10, (B, N), generator=g).cuda()
i_gt = torch.randint(-1, 12, (B, N), generator=g).cuda()
off = (torch.randn(B, N, 3, generator=g) * 0.05).cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    net.zero_grad(set_to_none=True)
    o = net(pts, nrm)
    (nll_loss(o["type_per_point"], t_gt) + 10 * offset_l1_loss(o["pt_offsets"], i_gt, off)).backward()
torch.cuda.synchronize()
