"""Host time to enqueue one fwd+bwd step vs. its device time (is the launching thread ahead of the GPU?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gcanet_b200 as gb
from gcanet_b200.synth import abc_like_batch

dev = torch.device("cuda", 0)
torch.manual_seed(0)
enc = gb.DGCNNEncoderGn(mode=0, nn_nb=50, input_channels=6).to(dev)
x = torch.from_numpy(abc_like_batch(16, 10000, seed=1234)).to(dev)
cot = [torch.randn(16, c, 10000, device=dev) for c in (64, 64, 128)]


def step():
    for p in enc.parameters():
        p.grad = None
    outs = enc.edge_stack(x)
    torch.autograd.backward(outs, cot)


for _ in range(5):
    step()
torch.cuda.synchronize()
K = 20
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
for _ in range(K):
    step()
b.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0) / K:.3f} ms/step, device {a.elapsed_time(b) / K:.3f} ms/step, "
      f"wall incl. drain {1e3 * (t2 - t0) / K:.3f} ms/step")
# one step enqueued onto an idle GPU: how long the first kernels wait for the host
torch.cuda.synchronize()
t0 = time.perf_counter()
step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"single step: host enqueue {1e3 * (t1 - t0):.3f} ms, until done {1e3 * (t2 - t0):.3f} ms")
