"""Input side of the hot path: raw ABC samples -> the batch the model and its losses consume.

The reference reads one ``.h5`` per shape (datasets ``points``, ``normals``, ``labels``, ``prim``, ``T_param``;
dataloader/ABCDataset_new.py:57-64) in 16 DataLoader workers, post-processes each sample with numpy
(:77-141, getInstanceInfo :157-178) and stacks the batch on the host (collate_fn :182-295).  Here the raw samples of a
shard live in device memory and ``prepare_batch`` produces a step's batch with ONE kernel launch, directly in the two
layouts the kernels want (channel-major [B, 6, N] and point-major [B, N, 8]); with several GPUs every rank holds and
prepares only its own clouds (``gcanet_b200.parallel.shard_range``).

The on-disk format is the reference's: ``read_h5_sample`` reads one of its files when ``h5py`` is importable (it is not in
the build image, hence untested there); ``save_shard`` / ``load_shard`` keep a set of samples as one ``.npz`` with the same
five arrays stacked, which needs nothing but numpy.
"""
from __future__ import annotations

import ctypes as _ct
from typing import Dict, Sequence

import numpy as np
import torch

from . import _cabi
from ._cabi import PrepareDesc, call, ptr, require_cuda, stream

RAW_KEYS = ("points", "normals", "labels", "prim", "T_param")


def read_h5_sample(path: str) -> Dict[str, np.ndarray]:
    """One sample of the reference's dataset (ABCDataset_new.py:57-64)."""
    try:
        import h5py
    except ImportError as exc:                                    # pragma: no cover - h5py is absent from the build image
        raise RuntimeError("h5py is not installed: convert the dataset with save_shard on a machine that has it") from exc
    with h5py.File(path, "r") as hf:
        return {"points": np.array(hf.get("points")), "normals": np.array(hf.get("normals")), "labels": np.array(hf.get("labels")),
                "prim": np.array(hf.get("prim")), "T_param": np.array(hf.get("T_param"))}


def save_shard(path: str, samples: Sequence[Dict[str, np.ndarray]]) -> None:
    np.savez_compressed(path, **{k: np.stack([np.asarray(s[k]) for s in samples]) for k in RAW_KEYS})


def load_shard(path: str, device) -> Dict[str, torch.Tensor]:
    """A shard of raw samples -> device tensors (points / normals / T_param fp32, labels / prim int32)."""
    z = np.load(path)
    out = {}
    for k in RAW_KEYS:
        a = z[k]
        t = torch.from_numpy(a.astype(np.int32) if k in ("labels", "prim") else a.astype(np.float32))
        out[k] = t.to(device).contiguous()
    return out


def draw_subsample(num_clouds: int, n_raw: int, n_sub: int = 7000, rng=None) -> np.ndarray:
    """The reference's subsample, one draw per cloud: ``np.random.choice(range(n_raw), n_sub, replace=False)`` (:120)."""
    rng = np.random if rng is None else rng
    return np.stack([rng.choice(range(n_raw), n_sub, replace=False) for _ in range(num_clouds)]).astype(np.int32)


def prepare_batch(raw: Dict[str, torch.Tensor], sub_index, num_primitives: int = 10, min_points: int = 100,
                  max_labels: int = 1024) -> Dict[str, torch.Tensor]:
    """raw: device tensors of B samples (``load_shard`` or a slice of it); sub_index [B, n_sub] int32 (host or device).
    Returns the reference's batch keys (``gt_pc``, ``gt_normal``, ``T_gt``, ``T_param``, ``I_gt``, ``I_gt_clean``,
    ``pt_offset_label``, ``instance_pointnum``, ``instance_cl``, ``batch_idx``; the voxel maps of the spconv head are not
    built) plus ``cloud_cn`` [B, 6, n] and ``cloud_nc`` [B, n, 8] for the kernels, and ``inst_num`` [B]."""
    pts = raw["points"]
    require_cuda(pts, "points", torch.float32)
    dev = pts.device
    sub = torch.as_tensor(sub_index, dtype=torch.int32).to(dev).contiguous()
    B, n_raw, _ = pts.shape
    n_sub = sub.shape[1]
    for k, dt in (("normals", torch.float32), ("T_param", torch.float32), ("labels", torch.int32), ("prim", torch.int32)):
        require_cuda(raw[k], k, dt)
    desc = PrepareDesc(B, n_raw, n_sub, max_labels, min_points, num_primitives)
    with torch.cuda.device(dev):
        f32 = dict(dtype=torch.float32, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        cloud_cn, cloud_nc = torch.empty((B, 6, n_sub), **f32), torch.empty((B, n_sub, 8), **f32)
        i_gt, t_gt, i_clean = (torch.empty((B, n_sub), **i32) for _ in range(3))
        t_param, off = torch.empty((B, n_sub, 22), **f32), torch.empty((B, n_sub, 3), **f32)
        inst_num = torch.empty(B, **i32)
        inst_pn, inst_cls = torch.empty((B, max_labels), **i32), torch.empty((B, max_labels), **i32)
        status = torch.empty(B, **i32)
        call("gcanet_prepare_samples", _ct.byref(desc), ptr(pts), ptr(raw["normals"]), ptr(raw["labels"]), ptr(raw["prim"]),
             ptr(raw["T_param"]), ptr(sub), ptr(cloud_cn), ptr(cloud_nc), ptr(i_gt), ptr(t_gt), ptr(i_clean), ptr(t_param), ptr(off),
             ptr(inst_num), ptr(inst_pn), ptr(inst_cls), ptr(status), stream())
    if int(status.max()) != 0:
        raise RuntimeError(f"prepare_batch: an instance label outside [0, {max_labels}) (raise max_labels)")
    counts = inst_num.tolist()
    keep = torch.arange(max_labels, device=dev).unsqueeze(0) < inst_num.unsqueeze(1)
    return {"gt_pc": cloud_nc[:, :, 0:3], "gt_normal": cloud_nc[:, :, 3:6], "T_gt": t_gt.long(), "T_param": t_param, "I_gt": i_gt,
            "I_gt_clean": i_clean, "pt_offset_label": off, "instance_pointnum": inst_pn[keep].long(), "instance_cl": inst_cls[keep].long(),
            "batch_idx": torch.arange(B, device=dev, dtype=torch.int32).repeat_interleave(n_sub), "inst_num": inst_num,
            "cloud_cn": cloud_cn, "cloud_nc": cloud_nc, "instances_per_cloud": counts}
