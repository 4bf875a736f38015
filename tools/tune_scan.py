"""Pass-A refresh schedule / unconditional prefix of the pruned scan (--aids builds: GCANET_TC_REFRESH, GCANET_TC_PRE),
xyz clouds and layer-1 activations, B = 16 x 10k, k = 50, set-only lists.  Prints ms per call for each setting."""
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
feats = {'xyz': x, 'x1': x1.contiguous(), 'x2': x2.contiguous()}

def t(f, n=20):
    for _ in range(3):
        G.knn_graph(f, k, k, want64=False, want32=True, ordered=False)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        G.knn_graph(f, k, k, want64=False, want32=True, ordered=False)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

settings = sys.argv[1:] or ["default", "4,2,32", "6,2,24", "6,3,18", "8,2,32", "8,4,8", "12,3,36", "4,2,64", "3,2,48"]
for name, f in feats.items():
    os.environ["GCANET_TC_STATS"] = "1"
    os.environ.pop("GCANET_TC_REFRESH", None)
    G.knn_graph(f, k, k, want64=False, want32=True, ordered=False)
    torch.cuda.synchronize()
    os.environ.pop("GCANET_TC_STATS")
    for sset in settings:
        if sset == "default":
            os.environ.pop("GCANET_TC_REFRESH", None)
        else:
            os.environ["GCANET_TC_REFRESH"] = sset
        print(f"{name} refresh={sset}: {t(f):.3f} ms per call", flush=True)
    os.environ.pop("GCANET_TC_REFRESH", None)
    for pre in ("4", "8", "12"):
        os.environ["GCANET_TC_PRE"] = pre
        print(f"{name} pre={pre}: {t(f):.3f} ms per call", flush=True)
    os.environ.pop("GCANET_TC_PRE", None)
