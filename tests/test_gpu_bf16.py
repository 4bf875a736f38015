"""GPU suite, part 5: the bf16-storage mode of the fused EdgeConv ("bf16 activations", BASELINE configs[1]).

In this mode the projected operand [P|Q] is rounded to bf16 where the projection GEMM writes it; the gather, GroupNorm
statistics, max, and every gradient are fp32.  Stated tolerance (SURVEY 8(c) rule 3): forward within 2e-2 of the
activation scale (bf16 has 8 mantissa bits: each of P_j and Q_i carries 2^-9 relative error; measured: 3e-3 .. 7e-3).
Gradients: neighbours whose pre-norm activations agree to bf16 rounding swap the arg-max, so gradient moves between
points and single rows of dW shift: median error below 1e-2 of the largest entry, single entries below 0.25, dX within
0.25 in relative L2 and conserved (sum over the points within 5e-3).  fp32 stays the parity mode and the default.
"""
import ctypes

import numpy as np
import pytest
import torch

import gcanet_b200 as gb
from gcanet_b200 import _cabi, functional as G
from gcanet_b200.synth import abc_like_batch
from oracle import dgcnn_oracle as orc

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel(a, b):
    e = (a.detach().cpu().double() - b.detach().double()).abs() / float(b.detach().abs().max())
    return float(e.max()), float(e.median())


@pytest.mark.parametrize("C,Cout,N,k", [(3, 64, 2000, 20), (64, 64, 3000, 50), (64, 128, 3000, 50), (6, 64, 1500, 30)])
def test_edgeconv_bf16_storage_vs_oracle(C, Cout, N, k):
    g = torch.Generator().manual_seed(C * 31 + Cout)
    B = 2
    if C <= 6:
        x = torch.from_numpy(abc_like_batch(B, N, seed=8, with_normals=(C == 6)))
    else:
        x = torch.randn(B, C, N, generator=g)
    W = torch.randn(Cout, 2 * C, generator=g) / (2 * C) ** 0.5
    gamma = torch.randn(Cout, generator=g) * 0.7 + 0.2
    beta = torch.randn(Cout, generator=g) * 0.3
    idx = orc.knn_points_normals(x, k, k) if C == 6 else orc.knn(x, k, k)
    cot = torch.randn(B, Cout, N, generator=g)
    xo, Wo, go, bo = (t.clone().requires_grad_(True) for t in (x, W, gamma, beta))
    out_o = orc.edgeconv_block(orc.get_graph_feature(xo, k, k, idx=idx), Wo, go, bo, groups=2)
    (out_o * cot).sum().backward()
    res = {}
    for storage in ("fp32", "bf16"):
        xg, Wg, gg, bg = (t.to(DEV).requires_grad_(True) for t in (x, W, gamma, beta))
        out_nc, out_cn = gb.edgeconv(G._ToPointMajor.apply(xg, (C + 3) // 4 * 4), idx.int().to(DEV), Wg, gg, bg, C, groups=2,
                                     storage=storage)
        (out_cn * cot.to(DEV)).sum().backward()
        res[storage] = (out_cn, xg.grad, Wg.grad, gg.grad, bg.grad)
    f32, b16 = res["fp32"], res["bf16"]
    mx, med = _rel(b16[0], out_o)
    print(f"[{C}->{Cout}] bf16 forward: max {mx:.2e}, median {med:.2e} (fp32 mode: max {_rel(f32[0], out_o)[0]:.2e})")
    assert mx < 2e-2 and med < 2e-3
    assert _rel(f32[0], out_o)[0] < 2e-4                      # the default mode is untouched
    assert not torch.equal(b16[0], f32[0])                     # and the bf16 mode really rounds
    for name, got, want in (("dW", b16[2], Wo.grad), ("dgamma", b16[3], go.grad), ("dbeta", b16[4], bo.grad)):
        mx, med = _rel(got, want)
        print(f"[{C}->{Cout}] bf16 {name}: max {mx:.2e}, median {med:.2e}")
        assert mx < 0.25 and med < 1e-2, name
    l2 = float((b16[1].cpu() - xo.grad).norm() / xo.grad.norm())
    print(f"[{C}->{Cout}] bf16 dx: relative L2 error {l2:.2e}")
    assert l2 < 0.25            # arg-max moves between neighbours whose y agree to bf16: gradient rows move, nothing is lost
    lost = (b16[1].cpu() - xo.grad).sum(dim=2).abs() / xo.grad.abs().sum(dim=2)
    assert float(lost.max()) < 5e-3


def test_bf16_storage_halves_the_saved_operand_and_runs_the_stack():
    L = _cabi.lib()
    d32 = _cabi.EdgeConvDesc(16, 10000, 64, 64, 128, 50, 2, 1e-5, 0.2, 0)
    d16 = _cabi.EdgeConvDesc(16, 10000, 64, 64, 128, 50, 2, 1e-5, 0.2, 1)
    s32, s16 = L.gcanet_edgeconv_saved_bytes(ctypes.byref(d32)), L.gcanet_edgeconv_saved_bytes(ctypes.byref(d16))
    pq32 = 16 * 10000 * 256 * 4
    assert s32 - s16 == pq32 // 2
    # the three-layer stack in bf16 storage against the oracle on the same neighbour lists
    B, N, k = 2, 10000, 50
    torch.manual_seed(0)
    ref = orc.DGCNNEncoderGn(mode=0, nn_nb=k, input_channels=6)
    enc = gb.DGCNNEncoderGn(mode=0, nn_nb=k, input_channels=6)
    enc.load_state_dict(ref.state_dict())
    enc.to(DEV)
    enc.storage = "bf16"
    enc.keep_graphs = True
    x = torch.from_numpy(abc_like_batch(B, N, seed=4321))
    outs = enc.edge_stack(x.to(DEV))
    graphs = [t.long().cpu() for t in enc.last_graphs]
    with torch.no_grad():
        x1 = ref.conv1(orc.get_graph_feature(x, k, k, idx=graphs[0])).max(dim=-1)[0]
        # layers 2 and 3 are compared on OUR inputs: their inputs already carry the bf16 noise of the layer below
        x2 = ref.conv2(orc.get_graph_feature(outs[0].detach().cpu(), k, k, idx=graphs[1])).max(dim=-1)[0]
        x3 = ref.conv3(orc.get_graph_feature(outs[1].detach().cpu(), k, k, idx=graphs[2])).max(dim=-1)[0]
    for name, a, b in zip(("x1", "x2", "x3"), outs, (x1, x2, x3)):
        mx, med = _rel(a, b)
        print(f"stack bf16 {name}: max {mx:.2e}, median {med:.2e}")
        assert mx < 2e-2 and med < 2e-3
    torch.autograd.backward(outs, [torch.randn_like(o) for o in outs])
    for name, p in enc.named_parameters():
        if name.split(".")[0] in ("conv1", "conv2", "conv3"):
            assert p.grad is not None and bool(torch.isfinite(p.grad).all())
