import sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
from gcanet_b200 import functional as G
from gcanet_b200.synth import abc_like_batch
for (N,k) in [(1001,50),(1001,20),(2000,50),(1001,41)]:
    x = torch.from_numpy(abc_like_batch(2, N, seed=100 + N)).cuda()
    a = G.knn_graph(x, k, k)[0]; b = G.knn_graph(x, k, k, brute_force=True)[0]
    srt = a.sort(dim=2)[0]; dup = (srt[:, :, 1:] == srt[:, :, :-1]).any(dim=2)
    srtb = b.sort(dim=2)[0]; dupb = (srtb[:, :, 1:] == srtb[:, :, :-1]).any(dim=2)
    print(N, k, 'equal', bool(torch.equal(a, b)), 'dup rows fast', int(dup.sum()), 'dup rows brute', int(dupb.sum()), 'mismatch rows', int((a != b).any(dim=2).sum()))
    if dup.any():
        bi, qi = [int(v[0]) for v in torch.nonzero(dup, as_tuple=True)]
        print(' row', bi, qi, a[bi, qi].tolist()); print(' brute', b[bi, qi].tolist())
