"""Forward / backward time of the offset-prediction block at the config-4 shape (B=16 x 10k, S=120, k=30, E=64)."""
import sys, torch
sys.path.insert(0, '/root/repo')
import gcanet_b200 as gb
torch.manual_seed(0)
B, N = 16, 10000
m = gb.OFFSET_PRED_MODULE(nn_nb=30, sampling_ratio=120).cuda()
pts = torch.rand(B, N, 3, device='cuda')
feat = torch.randn(B, N, 128, device='cuda', requires_grad=True)
inst = torch.randn(B, N, 64, device='cuda', requires_grad=True)
def run():
    out = m(pts, feat, inst)
    out.square().mean().backward()
for _ in range(3): run()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
tf = tb = 0.0
for _ in range(10):
    ev[0].record(); out = m(pts, feat, inst); ev[1].record(); out.square().mean().backward(); ev[2].record()
    torch.cuda.synchronize()
    tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2])
print(f"offset block: forward {tf / 10:.3f} ms, backward {tb / 10:.3f} ms")
