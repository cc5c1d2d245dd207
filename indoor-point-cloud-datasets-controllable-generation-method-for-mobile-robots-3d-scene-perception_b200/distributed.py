"""
Pose-sharded trajectories over the GPUs of one box.

Frames are independent given the mesh (the reference's loop carries no state, s3dis_simulator.py:254-288), so
the trajectory is cut into contiguous pose slices, one per rank; the mesh is replicated and every rank builds
the same deterministic LBVH locally.  The only exchange step is the final collection of the compacted clouds:
an all-gather of per-frame point counts followed by an all-gather of the records, padded to the largest
per-rank count (NCCL has no all-gatherv) and trimmed on arrival.  Noise is keyed on the GLOBAL pose index, so
the gathered result is bit-identical to a single-GPU run.

One process per GPU (``torchrun``); ``torch.distributed`` is plumbing only.  The gather logic is backend-
agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist

from .trajectory import shard_range

RECORD_KEYS = ("points", "incident", "prim_id", "label", "ray_idx")


@dataclass
class GatheredCloud:
    """Whole-trajectory result, identical on every rank.  Tensors live where the local results lived."""
    points: torch.Tensor
    incident: torch.Tensor
    prim_id: torch.Tensor
    label: torch.Tensor
    ray_idx: torch.Tensor
    frame_offset: torch.Tensor      # (P_total + 1,) int64

    def numpy(self) -> Dict[str, np.ndarray]:
        out = {k: getattr(self, k).cpu().numpy() for k in RECORD_KEYS}
        for k in ("prim_id", "label", "ray_idx"):
            out[k] = out[k].view(np.uint32)
        out["frame_offset"] = self.frame_offset.cpu().numpy()
        return out


def world_info(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def allgather_clouds(local: Dict[str, torch.Tensor], frame_counts: torch.Tensor, num_poses_total: int, group=None,
                     keys=RECORD_KEYS) -> GatheredCloud:
    """Collect every rank's compacted records in rank (= pose) order.

    local         tensors of this rank, first dimension = number of local points (ray-ordered, frame-major)
    frame_counts  (P_local,) int64: points per local frame
    """
    rank, world = world_info(group)
    dev = frame_counts.device
    shard_sizes = [len(shard_range(num_poses_total, r, world)) for r in range(world)]
    if world == 1:
        off = torch.zeros(num_poses_total + 1, dtype=torch.int64, device=dev)
        off[1:] = torch.cumsum(frame_counts, 0)
        vals = {k: local[k] for k in keys}
        return GatheredCloud(frame_offset=off, **{k: vals.get(k) for k in RECORD_KEYS})
    # 1) per-frame counts, padded to the largest shard
    pmax = max(shard_sizes)
    mine = torch.zeros(pmax, dtype=torch.int64, device=dev)
    mine[: frame_counts.numel()] = frame_counts
    allc = torch.empty(world * pmax, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, mine, group=group)
    allc = allc.view(world, pmax)
    counts = torch.cat([allc[r, : shard_sizes[r]] for r in range(world)])
    off = torch.zeros(num_poses_total + 1, dtype=torch.int64, device=dev)
    off[1:] = torch.cumsum(counts, 0)
    per_rank = [int(x) for x in allc.sum(dim=1).tolist()]          # one small D2H: sizes of the trimmed slices
    mmax = max(per_rank)
    # 2) records, padded to the largest per-rank point count
    gathered = {}
    for k in keys:
        src = local[k]
        pad = torch.zeros((mmax,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        pad[: src.shape[0]] = src
        dst = torch.empty((world * mmax,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
        dist.all_gather_into_tensor(dst, pad, group=group)
        dst = dst.view((world, mmax) + tuple(src.shape[1:]))
        gathered[k] = torch.cat([dst[r, : per_rank[r]] for r in range(world)])
    return GatheredCloud(frame_offset=off, **{k: gathered.get(k) for k in RECORD_KEYS})


def simulate_sharded(engine, poses_all, intrinsics, mesh=None, noise=None, group=None) -> GatheredCloud:
    """Run ``engine.simulate`` on this rank's contiguous slice of ``poses_all`` and all-gather the clouds.

    Every rank passes the SAME ``poses_all`` (P,4,4) and mesh.  ``noise.pose_index_base`` is offset by the slice
    start so that frame p draws the same Philox stream whatever the number of ranks."""
    from .core import NoiseConfig
    rank, world = world_info(group)
    poses_all = np.ascontiguousarray(poses_all, dtype=np.float64).reshape(-1, 4, 4)
    sl = shard_range(len(poses_all), rank, world)
    local_noise = None
    if noise is not None:
        local_noise = NoiseConfig(noise.angle_noise_std, noise.dropout_probability, noise.range_noise_std, noise.seed,
                                  noise.pose_index_base + sl.start)
    res = engine.simulate(poses_all[sl.start:sl.stop], intrinsics, mesh, noise=local_noise)
    local = {"points": res.points, "incident": res.incident, "prim_id": res.prim_id, "label": res.label, "ray_idx": res.ray_idx}
    counts = (res.frame_offset[1:] - res.frame_offset[:-1]).contiguous()
    return allgather_clouds(local, counts, len(poses_all), group=group)
