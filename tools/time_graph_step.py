"""A/B: the benchmark step eager vs captured in one CUDA graph (forward + backward of the three-layer stack)."""
import statistics, sys, torch
sys.path.insert(0, '/root/repo')
import gcanet_b200 as gb
from gcanet_b200.synth import abc_like_batch
torch.manual_seed(0)
B, N, k = 16, 10000, 50
enc = gb.DGCNNEncoderGn(mode=0, nn_nb=k, input_channels=6).cuda()
hot = [p for n, p in enc.named_parameters() if n.split(".")[0] in ("conv1", "conv2", "conv3", "bn1", "bn2", "bn3")]
x = torch.from_numpy(abc_like_batch(B, N, seed=1234)).cuda()
cot = [torch.randn(B, c, N, device="cuda") for c in (64, 64, 128)]

def step(xin):
    for p in hot:
        p.grad = None
    outs = enc.edge_stack(xin)
    torch.autograd.backward(outs, cot)
    return outs[2].detach().sum()

def med(fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)

print("eager", round(med(lambda: step(x)), 4))
static_x = x.clone()
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): step(static_x)
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    loss = step(static_x)
g.replay(); torch.cuda.synchronize()
ref_loss = float(step(x)); g.replay(); torch.cuda.synchronize()
print("loss eager/graph", ref_loss, float(loss))
gw = [p.grad.clone() for p in hot]
step(x); torch.cuda.synchronize()
print("grad max diff", max(float((a - p.grad).abs().max()) for a, p in zip(gw, hot)))
print("graph", round(med(lambda: g.replay()), 4))
