import sys, torch
sys.path.insert(0, '/root/repo')
from gcanet_b200 import functional as G
torch.manual_seed(0)
x = torch.randn(16, 64, 10000, device='cuda')
import os
prune = os.environ.get('PR','1') == '1'
for _ in range(3): G.knn_graph(x, 50, 50, want64=False, want32=True, prune=prune)
torch.cuda.synchronize()
