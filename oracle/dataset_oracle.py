"""CPU oracle for the input side of the hot path: what ``ABCDataset.__getitem__`` does to one raw
sample after reading it, and ``getInstanceInfo`` (dataloader/ABCDataset_new.py:77-141, 157-178).  TEST INFRASTRUCTURE ONLY.

Restated with numpy:
  * instances with at most 100 raw points become background (-1); the kept ones are renumbered in order of FIRST
    APPEARANCE in the label array (``Counter`` iterates in insertion order, :84-89);
  * ``T_gt`` = primitive type where the instance survived, else -1; with 7 primitive classes 7 -> 6, 9 -> 6, 8 -> 2 (:91-98);
  * ``I_gt_clean`` = new id for kept instances, ``old label + number of kept instances`` for the small ones (:106-110);
  * subsample of 7000 points drawn WITHOUT replacement (:120-126; the caller passes the indices);
  * ``getInstanceInfo``: per kept instance the mean of its (subsampled) points, its size, the type of its first point;
    ``pt_offset_label`` = mean - point, with -100 standing in for the mean of background points (:157-178).

Pinned by oracle/make_golden.py, which executes the reference's own text for these lines on a synthetic raw sample and
asserts bit-identical outputs.
"""
from __future__ import annotations

import numpy as np


def prepare_sample(points, normals, labels, primitives, t_param, subidx, num_primitives=10, min_points=100):
    """One raw sample (points/normals [N, 3] fp32, labels/primitives [N] int, t_param [N, 22]) + subsample indices ->
    dict with the reference's keys."""
    labels = np.asarray(labels)
    counts = np.bincount(labels, minlength=int(labels.max()) + 1)
    seen, order = set(), []
    for v in labels.tolist():                                  # first-appearance order of the labels
        if v not in seen:
            seen.add(v)
            order.append(v)
    keys = [k for k in order if counts[k] > min_points]
    mapper = -np.ones(int(labels.max()) + 1)
    if keys:
        mapper[keys] = np.arange(len(keys))
    inst = mapper[labels]
    i_gt = inst.astype(int)
    clean = -np.ones_like(primitives)
    valid = inst != -1
    clean[valid] = primitives[valid]
    if num_primitives == 7:
        clean[clean == 7] = 6
        clean[clean == 9] = 6
        clean[clean == 8] = 2
    i_clean = inst.copy()
    small = inst == -1
    i_clean[small] = labels[small] + len(keys)
    out = {"gt_pc": points[subidx], "gt_normal": normals[subidx], "T_gt": clean.astype(int)[subidx], "T_param": t_param[subidx],
           "I_gt": i_gt[subidx], "I_gt_clean": i_clean.astype(int)[subidx]}
    num, pointnum, cls, off = instance_info(out["gt_pc"], out["I_gt"].astype(np.int32), out["T_gt"])
    out.update(inst_num=num, inst_pointnum=pointnum, inst_cls=cls, pt_offset_label=off)
    return out


def instance_info(xyz, instance_label, semantic_label):
    mean = np.full((xyz.shape[0], 3), -100.0, dtype=np.float32)
    num = max(int(instance_label.max()) + 1, 0)
    pointnum, cls = [], []
    for i in range(num):
        members = np.where(instance_label == i)
        mean[members] = xyz[members].mean(0)
        pointnum.append(members[0].size)
        cls.append(semantic_label[members[0][0]])
    return num, pointnum, cls, mean - xyz


def synthetic_raw_sample(n=8000, seed=0):
    """A raw ABC-like sample: points on primitive patches with instance labels (some instances smaller than 100 points),
    primitive types 0..9 and 22 parameters per point."""
    from gcanet_b200.synth import abc_like_batch
    rs = np.random.RandomState(seed)
    c = abc_like_batch(1, n, seed=seed + 500, with_normals=True)[0]          # [6, n]
    pts, nrm = np.ascontiguousarray(c[:3].T), np.ascontiguousarray(c[3:].T)
    # instances: k-means-like partition by a few random centres, plus a handful of tiny instances
    centres = pts[rs.choice(n, 9, replace=False)]
    lab = np.argmin(((pts[:, None] - centres[None]) ** 2).sum(-1), axis=1).astype(np.int64)
    tiny = rs.choice(n, 180, replace=False)
    lab[tiny[:60]] = 9
    lab[tiny[60:110]] = 10
    lab[tiny[110:]] = 12                                                       # label 11 never occurs
    lab = rs.permutation(13)[lab]                                              # labels in no particular order
    prim_of = rs.randint(0, 10, 13)
    prim = prim_of[lab].astype(np.int64)
    t_param = rs.randn(n, 22).astype(np.float32)
    return pts.astype(np.float32), nrm.astype(np.float32), lab, prim, t_param
