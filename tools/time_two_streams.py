"""A/B: the benchmark step on the full batch (one stream) vs two half-batches on two streams (kNN of one half overlapping the
EdgeConv of the other), eager and as one CUDA graph with two branches."""
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
halves = [(x[:8].contiguous(), [c[:8].contiguous() for c in cot]), (x[8:].contiguous(), [c[8:].contiguous() for c in cot])]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def full():
    for p in hot: p.grad = None
    torch.autograd.backward(enc.edge_stack(x), cot)

def two():
    for p in hot: p.grad = None
    cur = torch.cuda.current_stream()
    outs = []
    for s, (xh, ch) in zip((s1, s2), halves):
        s.wait_stream(cur)
        with torch.cuda.stream(s):
            outs.append((enc.edge_stack(xh), ch))
    for s, (o, ch) in zip((s1, s2), outs):
        with torch.cuda.stream(s):
            torch.autograd.backward(o, ch)
    cur.wait_stream(s1); cur.wait_stream(s2)

def med(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)

print("full eager", round(med(full), 4))
print("two streams eager", round(med(two), 4))
g_ref = None
for name, fn in (("full", full), ("two", two)):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2): fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            fn()
        g.replay(); torch.cuda.synchronize()
        gr = torch.cat([p.grad.reshape(-1) for p in hot]).clone()
        if g_ref is None: g_ref = gr
        else: print("grad rel diff two vs full", float((gr - g_ref).abs().max() / g_ref.abs().max()))
        print(name, "graph", round(med(lambda: g.replay()), 4))
    except Exception as e:
        print(name, "graph capture failed:", type(e).__name__, str(e)[:200])
        torch.cuda.synchronize()
