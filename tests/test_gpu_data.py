"""GPU suite, part 7: the input side (dataloader/ABCDataset_new.py:77-141, 157-178, collate :182-295) -- batch preparation
on the device against the oracle, which is bit-identical to the reference's own text (oracle/make_golden.py).

Integer outputs (instance ids, primitive types, sizes, per-instance class) and gathered arrays are bit-exact; the offset
labels agree to 2e-6 (the reference's mean is a sequential fp32 sum, the kernel's an fp64 sum rounded once).
"""
import os

import numpy as np
import pytest
import torch

from gcanet_b200 import data as D
from oracle import dataset_oracle as dso

DEV = "cuda"


@pytest.mark.gpu
@pytest.mark.parametrize("num_primitives", [10, 7])
def test_prepare_batch_vs_oracle(tmp_path, num_primitives):
    B, n_raw, n_sub = 3, 8000, 7000
    samples = []
    for b in range(B):
        p, n, l, pr, tp = dso.synthetic_raw_sample(n_raw, seed=20 + b)
        samples.append({"points": p, "normals": n, "labels": l, "prim": pr, "T_param": tp})
    shard = os.path.join(tmp_path, "shard.npz")
    D.save_shard(shard, samples)
    raw = D.load_shard(shard, DEV)
    sub = D.draw_subsample(B, n_raw, n_sub, rng=np.random.RandomState(5))
    batch = D.prepare_batch(raw, sub, num_primitives=num_primitives)
    assert batch["cloud_cn"].shape == (B, 6, n_sub) and batch["cloud_nc"].shape == (B, n_sub, 8)
    assert torch.equal(batch["cloud_cn"].transpose(1, 2), batch["cloud_nc"][:, :, :6]) and not batch["cloud_nc"][:, :, 6:].any()
    pn, cl = [], []
    for b in range(B):
        s = samples[b]
        want = dso.prepare_sample(s["points"], s["normals"], s["labels"], s["prim"], s["T_param"], sub[b], num_primitives=num_primitives)
        for key in ("gt_pc", "gt_normal", "T_param"):
            assert np.array_equal(batch[key][b].cpu().numpy(), want[key]), key
        for key in ("T_gt", "I_gt", "I_gt_clean"):
            assert np.array_equal(batch[key][b].cpu().numpy().astype(np.int64), np.asarray(want[key]).astype(np.int64)), key
        assert float(np.abs(batch["pt_offset_label"][b].cpu().numpy() - want["pt_offset_label"]).max()) < 2e-6
        assert int(batch["inst_num"][b]) == want["inst_num"]
        pn += list(want["inst_pointnum"])
        cl += [int(v) for v in want["inst_cls"]]
    assert batch["instance_pointnum"].tolist() == pn and batch["instance_cl"].tolist() == cl     # collate's extend() order
    assert batch["batch_idx"].shape == (B * n_sub,) and int(batch["batch_idx"][n_sub]) == 1
    # feeds the kernels directly: the stack accepts cloud_cn
    import gcanet_b200 as gb
    enc = gb.DGCNNEncoderGn(mode=5, nn_nb=20, input_channels=6).to(DEV)
    with torch.no_grad():
        x1, _, _ = enc.edge_stack(batch["cloud_cn"])
    assert x1.shape == (B, 64, n_sub)


@pytest.mark.gpu
def test_prepare_batch_rejects_labels_out_of_range():
    p, n, l, pr, tp = dso.synthetic_raw_sample(2000, seed=1)
    raw = {"points": torch.from_numpy(p)[None].to(DEV), "normals": torch.from_numpy(n)[None].to(DEV),
           "labels": torch.from_numpy(l.astype(np.int32))[None].to(DEV), "prim": torch.from_numpy(pr.astype(np.int32))[None].to(DEV),
           "T_param": torch.from_numpy(tp)[None].to(DEV)}
    with pytest.raises(RuntimeError, match="max_labels"):
        D.prepare_batch(raw, np.arange(1000, dtype=np.int32)[None], max_labels=8)


def test_shard_round_trip_and_subsample_cpu(tmp_path):
    p, n, l, pr, tp = dso.synthetic_raw_sample(500, seed=3)
    path = os.path.join(tmp_path, "s.npz")
    D.save_shard(path, [{"points": p, "normals": n, "labels": l, "prim": pr, "T_param": tp}] * 2)
    z = np.load(path)
    assert z["points"].shape == (2, 500, 3) and z["T_param"].shape == (2, 500, 22) and np.array_equal(z["labels"][1], l)
    sub = D.draw_subsample(2, 500, 300, rng=np.random.RandomState(0))
    assert sub.shape == (2, 300) and all(len(set(r.tolist())) == 300 for r in sub)
    np.random.seed(9)
    a = D.draw_subsample(1, 500, 300)
    np.random.seed(9)
    assert np.array_equal(a[0], np.random.choice(range(500), 300, replace=False))          # the reference's draw (:120)
