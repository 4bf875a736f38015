"""xyz kNN -- the tensor-core scan (C = 3, 1 024 <= N < 32 768) and the CUDA-core pruned kernel (`prune=False`, points x
normals, other sizes) -- against the brute-force scan: the lists must be identical -- ordered lists element for element,
unordered lists as sets -- on smooth clouds, uniform noise, grids (ties) and duplicates; then timings of both.
Usage: python tools/check_xyz.py [quick]"""
import sys

import numpy as np
import torch

sys.path.insert(0, '/root/repo')
from gcanet_b200 import functional as G
from gcanet_b200.synth import abc_like_batch


def clouds(kind, B, N, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "abc":
        return torch.from_numpy(abc_like_batch(B, N, seed=seed, with_normals=True))
    if kind == "uniform":
        p = torch.rand(B, 3, N, generator=g) * 2 - 1
    elif kind == "grid":            # massive exact ties
        m = int(round(N ** (1 / 3))) + 1
        ax = torch.arange(m, dtype=torch.float32) / m
        gr = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), 0).reshape(3, -1)[:, :N]
        p = gr.unsqueeze(0).repeat(B, 1, 1)
    elif kind == "dup":             # every point four times
        q = torch.rand(B, 3, N // 4, generator=g)
        p = q.repeat(1, 1, 4)[:, :, torch.randperm(N, generator=g)]
    elif kind == "same":            # one location
        p = torch.ones(B, 3, N) * 0.3
    n = torch.randn(B, 3, N, generator=g)
    n = n / n.norm(dim=1, keepdim=True)
    return torch.cat([p, n], 1)


def check(kind, B, N, k1, k2, metric, ordered, seed=0):
    x = clouds(kind, B, N, seed).cuda()
    x = x if metric == G.METRIC_POINTS_NORMALS else x[:, :3].contiguous()
    a, _ = G.knn_graph(x, k1, k2, metric, ordered=ordered)
    b, _ = G.knn_graph(x, k1, k2, metric, brute_force=True, ordered=True)
    if ordered:
        bad = int((a != b).any(-1).sum())
    else:
        bad = int((a.sort(-1)[0] != b.sort(-1)[0]).any(-1).sum())
    tag = f"{kind:8s} B={B} N={N} k=({k1},{k2}) metric={metric} ordered={ordered}"
    print(("ok   " if bad == 0 else "FAIL ") + tag + (f": {bad} rows differ" if bad else ""), flush=True)
    return bad == 0


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


ok = True
for kind in ("abc", "uniform", "grid", "dup", "same"):
    for metric in (G.METRIC_L2, G.METRIC_POINTS_NORMALS):
        for (k1, k2, ordered) in ((50, 50, False), (50, 50, True), (20, 20, True), (64, 64, False), (10, 40, True), (1, 1, True)):
            ok &= check(kind, 2, 3000 if kind != "same" else 600, k1, k2, metric, ordered)
ok &= check("same", 2, 2048, 50, 50, G.METRIC_L2, False)       # every row overflows: the per-row fallback
ok &= check("grid", 2, 4096, 50, 50, G.METRIC_L2, True)
ok &= check("abc", 2, 5000, 80, 80, G.METRIC_L2, True)          # 64 < k <= 128: two slot minima per column
ok &= check("abc", 16, 10000, 50, 50, G.METRIC_L2, False)
ok &= check("abc", 16, 10000, 50, 50, G.METRIC_L2, True)
ok &= check("abc", 16, 10000, 50, 50, G.METRIC_POINTS_NORMALS, False)
ok &= check("abc", 3, 1037, 20, 20, G.METRIC_L2, True)
ok &= check("abc", 2, 257, 20, 20, G.METRIC_L2, True)
if len(sys.argv) < 2:
    ok &= check("abc", 2, 100000, 50, 50, G.METRIC_L2, False)
print("ALL OK" if ok else "FAILURES")

x6 = torch.from_numpy(abc_like_batch(16, 10000, seed=1, with_normals=True)).cuda()
x3 = x6[:, :3].contiguous()
for name, x, m in (("L2 C=3", x3, G.METRIC_L2), ("PN C=6", x6, G.METRIC_POINTS_NORMALS)):
    for k in (20, 50):
        t = timeit(lambda: G.knn_graph(x, k, k, m, want64=False, want32=True, ordered=False))
        t2 = timeit(lambda: G.knn_graph(x, k, k, m, want64=True, want32=False, ordered=True))
        t3 = timeit(lambda: G.knn_graph(x, k, k, m, want64=False, want32=True, ordered=False, prune=False))
        print(f"{name} k={k}: unordered int32 {t:.3f} ms, ordered int64 {t2:.3f} ms per call (B=16 x 10k); CUDA-core kernel (prune=False) {t3:.3f} ms")
xl = torch.from_numpy(abc_like_batch(4, 100000, seed=2, with_normals=False)).cuda()
print(f"L2 C=3 k=50 B=4 x 100k: {timeit(lambda: G.knn_graph(xl, 50, 50, G.METRIC_L2, want64=False, want32=True, ordered=False), 5):.3f} ms"
      f" (CUDA-core kernel: {timeit(lambda: G.knn_graph(xl, 50, 50, G.METRIC_L2, want64=False, want32=True, ordered=False, prune=False), 5):.3f} ms)")
