"""Seeded synthetic "ABC-shaped" point clouds.

The ABC/HPNet samples the reference trains on (dataloader/ABCDataset_new.py:57-64)
are CAD shapes: a handful of planes, cylinders, cones and spheres, 10 000 points
with unit normals, mean-centred and scaled so the largest bounding-box extent is 1
(utils/process_abc.py:49-73).  There is no dataset in this environment, so the
benchmarks and parity tests draw clouds with the same structure from a seeded
generator: cloud ``i`` of a run uses ``numpy.random.RandomState(seed + i)``.
"""
from __future__ import annotations

import numpy as np


def _rand_rotation(rs: np.random.RandomState) -> np.ndarray:
    q, r = np.linalg.qr(rs.randn(3, 3))
    q = q * np.sign(np.diag(r))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


def _plane(rs, n):
    w, h = rs.uniform(0.3, 1.0, 2)
    p = np.stack([rs.uniform(-w / 2, w / 2, n), rs.uniform(-h / 2, h / 2, n), np.zeros(n)], 1)
    nr = np.tile(np.array([0.0, 0.0, 1.0]), (n, 1))
    return p, nr, w * h


def _cylinder(rs, n):
    r, h = rs.uniform(0.1, 0.4), rs.uniform(0.3, 1.0)
    arc = rs.uniform(np.pi / 2, 2 * np.pi)
    t = rs.uniform(0, arc, n)
    z = rs.uniform(-h / 2, h / 2, n)
    nr = np.stack([np.cos(t), np.sin(t), np.zeros(n)], 1)
    return np.concatenate([r * nr[:, :2], z[:, None]], 1), nr, arc * r * h


def _sphere(rs, n):
    r = rs.uniform(0.15, 0.45)
    zmin = rs.uniform(-1.0, 0.5)
    z = rs.uniform(zmin, 1.0, n)
    t = rs.uniform(0, 2 * np.pi, n)
    s = np.sqrt(np.maximum(0.0, 1 - z * z))
    nr = np.stack([s * np.cos(t), s * np.sin(t), z], 1)
    return r * nr, nr, 2 * np.pi * r * r * (1.0 - zmin)


def _cone(rs, n):
    half = rs.uniform(np.pi / 12, np.pi / 4)
    l0, l1 = sorted(rs.uniform(0.1, 1.0, 2))
    l1 = max(l1, l0 + 0.1)
    # area-uniform along the slant length
    l = np.sqrt(rs.uniform(l0 * l0, l1 * l1, n))
    t = rs.uniform(0, 2 * np.pi, n)
    sa, ca = np.sin(half), np.cos(half)
    p = np.stack([l * sa * np.cos(t), l * sa * np.sin(t), l * ca], 1)
    nr = np.stack([ca * np.cos(t), ca * np.sin(t), -sa * np.ones(n)], 1)
    return p, nr, np.pi * sa * (l1 * l1 - l0 * l0)


_PRIMS = (_plane, _cylinder, _sphere, _cone)


def abc_like_cloud(num_points: int, seed: int, noise: float = 0.0):
    """Returns (points [N,3], normals [N,3]) float32; normals are unit length."""
    rs = np.random.RandomState(seed)
    n_prim = rs.randint(4, 13)
    kinds = rs.randint(0, len(_PRIMS), n_prim)
    # draw each patch once at unit density to learn its area, then apportion N
    metas = []
    for kd in kinds:
        sub = np.random.RandomState(rs.randint(0, 2 ** 31 - 1))
        state = sub.get_state()
        _, _, area = _PRIMS[kd](sub, 1)
        metas.append((kd, state, area, _rand_rotation(rs), rs.uniform(-0.5, 0.5, 3)))
    areas = np.array([m[2] for m in metas])
    counts = np.floor(areas / areas.sum() * num_points).astype(np.int64)
    counts[np.argmax(counts)] += num_points - counts.sum()
    pts, nrm = [], []
    for (kd, state, _, rot, shift), cnt in zip(metas, counts):
        if cnt <= 0:
            continue
        sub = np.random.RandomState()
        sub.set_state(state)
        p, nr, _ = _PRIMS[kd](sub, int(cnt))
        pts.append(p @ rot.T + shift)
        nrm.append(nr @ rot.T)
    p = np.concatenate(pts, 0)
    nr = np.concatenate(nrm, 0)
    if noise > 0:
        p = p + nr * np.clip(rs.randn(p.shape[0], 1) * noise, -noise, noise)
    perm = rs.permutation(p.shape[0])
    p, nr = p[perm], nr[perm]
    p = p - p.mean(0, keepdims=True)
    p = p / (p.max(0) - p.min(0)).max()
    nr = nr / np.linalg.norm(nr, axis=1, keepdims=True)
    return p.astype(np.float32), nr.astype(np.float32)


def abc_like_batch(batch: int, num_points: int, seed: int = 1234, with_normals: bool = False,
                   first_cloud: int = 0) -> np.ndarray:
    """Channel-major batch ``[B, 3 or 6, N]`` float32 (the layout the reference's
    encoder takes, train_new.py:25-26).  ``first_cloud`` offsets the per-cloud seed so
    ranks of a sharded run draw disjoint clouds."""
    out = np.empty((batch, 6 if with_normals else 3, num_points), np.float32)
    for b in range(batch):
        p, nr = abc_like_cloud(num_points, seed + first_cloud + b)
        out[b, 0:3] = p.T
        if with_normals:
            out[b, 3:6] = nr.T
    return out
