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


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, BEFORE any pinned host memory is allocated.
    One process per GPU is launched without any placement (torchrun); page-locked staging buffers that land on the
    other socket send every PCIe transfer across the inter-socket link.  Returns the node, or None when the topology
    cannot be read (nothing is changed then)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


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


class OverlappedShardedScan:
    """Steady-state multi-GPU trajectory scan with the cloud exchange hidden behind the traversal.

    This rank's pose slice is cut into ``chunks`` pieces.  Every chunk writes its compacted ``xyz | label |
    frame_offset`` into ONE contiguous device block of fixed capacity; as soon as a chunk's kernels are done an
    asynchronous all-gather of that block runs on the communicator's stream while the next chunk is traversed.  No
    host synchronisation happens inside ``step()`` (fixed-capacity blocks need no counts on the host; the gathered
    per-frame offsets say which part of every block is valid).

    Gathered layout, per chunk c: ``gathered[c]`` is a (world, block_bytes) uint8 tensor; ``views(c, r)`` returns the
    (xyz, label, frame_offset) views of rank r's chunk c.  Incident angles stay local: they are a pure function of a
    point and its frame pose (reference raycast_engine_cpu.py:100-107) and moving them would add 50 % link traffic.
    """

    def __init__(self, ctx, poses_local, intrinsics, noise=None, chunks: int = 4, group=None):
        from . import core
        self.ctx, self.intr, self.group = ctx, intrinsics, group
        self.rank, self.world = world_info(group)
        dev = ctx.device
        self.n_frame = core.rays_per_frame(intrinsics)
        poses_local = np.ascontiguousarray(poses_local, dtype=np.float64).reshape(-1, 16)
        self.P = len(poses_local)
        self.poses_d = torch.from_numpy(poses_local).to(dev)
        chunks = max(1, min(int(chunks), self.P)) if self.P > 0 else 1
        bounds = [(self.P * c) // chunks for c in range(chunks + 1)]
        self.ranges = [(bounds[c], bounds[c + 1]) for c in range(chunks) if bounds[c + 1] > bounds[c]]
        self.noise = noise
        self.comm = torch.cuda.Stream(dev)
        self.blocks, self.layout, self.bufs, self.gathered = [], [], [], []
        for (a, b) in self.ranges:
            nf, cap = b - a, (b - a) * self.n_frame
            o_lab = cap * 12
            o_off = (o_lab + cap * 4 + 255) // 256 * 256
            nbytes = (o_off + (nf + 1) * 8 + 255) // 256 * 256
            blk = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
            xyz = blk[: cap * 12].view(torch.float32).view(cap, 3)
            lab = blk[o_lab: o_lab + cap * 4].view(torch.int32)
            off = blk[o_off: o_off + (nf + 1) * 8].view(torch.int64)
            self.blocks.append(blk)
            self.layout.append((cap, nf, o_lab, o_off))
            self.bufs.append({"xyz": xyz, "incident": torch.empty(cap, dtype=torch.float64, device=dev), "prim": None,
                              "label": lab, "ray": None, "off": off})
            self.gathered.append(torch.empty((self.world, nbytes), dtype=torch.uint8, device=dev) if self.world > 1 else blk.view(1, -1))

    def step(self):
        """Enqueue one whole trajectory pass (+ exchange).  Returns after enqueueing; the caller's stream is made to
        wait for the exchange, so an event recorded afterwards covers everything."""
        from .core import NoiseConfig
        main = torch.cuda.current_stream(self.ctx.device)
        works = []
        for c, (a, b) in enumerate(self.ranges):
            nz = None
            if self.noise is not None:
                nz = NoiseConfig(self.noise.angle_noise_std, self.noise.dropout_probability, self.noise.range_noise_std,
                                 self.noise.seed, self.noise.pose_index_base + a)
            self.ctx.scan_enqueue(self.poses_d[a:b], self.intr, nz, self.bufs[c])
            if self.world > 1:
                ev = torch.cuda.Event()
                ev.record(main)
                with torch.cuda.stream(self.comm):
                    self.comm.wait_event(ev)
                    works.append(dist.all_gather_into_tensor(self.gathered[c].view(-1), self.blocks[c], group=self.group, async_op=True))
        for w in works:
            w.wait()          # the caller's current stream waits for the collective, the host does not
        return works

    def views(self, c: int, r: int):
        cap, nf, o_lab, o_off = self.layout[c]
        g = self.gathered[c][r]
        return (g[: cap * 12].view(torch.float32).view(cap, 3), g[o_lab: o_lab + cap * 4].view(torch.int32),
                g[o_off: o_off + (nf + 1) * 8].view(torch.int64))

    def assemble_numpy(self) -> Dict[str, np.ndarray]:
        """Dense, pose-ordered cloud of ALL ranks (testing / export; synchronises).  Requires every rank to have used
        the same chunking, which holds when shards have equal sizes."""
        torch.cuda.synchronize(self.ctx.device)
        pts, labs, counts = [], [], []
        for r in range(self.world):
            for c in range(len(self.ranges)):
                xyz, lab, off = self.views(c, r)
                o = off.cpu().numpy()
                m = int(o[-1])
                pts.append(xyz[:m].cpu().numpy()); labs.append(lab[:m].cpu().numpy().view(np.uint32))
                counts.append(np.diff(o))
        cnt = np.concatenate(counts) if counts else np.zeros(0, np.int64)
        return {"points": np.concatenate(pts), "label": np.concatenate(labs),
                "frame_offset": np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)}


class _RawCudaBuffer:
    """Exposes a raw device pointer through __cuda_array_interface__ so torch can view it without a copy."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class PeerGather:
    """All-gather of the compacted clouds over NVLink peer memory (``lrc_set_gather``).

    Every rank owns one buffer ``[xyz: world*cap x 3 f32 | label: world*cap u32 | frame_offset: world*(Pmax+1) i64]``
    allocated by the engine and mapped by all other ranks through CUDA IPC.  While enabled, every scan pushes rank r's
    compacted points into region r of ALL buffers (its own and the peers') with the library's exchange kernel (16-byte
    vector stores), chunk by chunk, while the next pose chunk is traversed -- there is no separate collective.  After ``synchronize()`` (stream sync
    + barrier) each rank holds the whole cloud: frame f of rank r is
    ``xyz[r*cap + off[r, f] - r*cap ...]`` -- see ``views`` / ``assemble_numpy``.
    """

    def __init__(self, ctx, cap_per_rank: int, frames_per_rank: int, group=None):
        import ctypes as C
        from . import _native as nat
        self.ctx, self.group = ctx, group
        self.rank, self.world = world_info(group)
        if self.world > nat.Gather.xyz.size // C.sizeof(C.c_void_p):
            raise ValueError("PeerGather supports at most 16 ranks")
        self.cap, self.pmax = (int(cap_per_rank) + 3) // 4 * 4, int(frames_per_rank)      # a multiple of 4 points: every rank's region keeps 16-byte alignment (TMA bulk copies)
        w = self.world
        self.o_lab = (w * self.cap * 12 + 255) // 256 * 256
        self.o_off = (self.o_lab + w * self.cap * 4 + 255) // 256 * 256
        # compact wire format (enable(wire=True)): hit distance and ray index per point, one progress word per rank
        self.o_t = (self.o_off + w * (self.pmax + 1) * 8 + 255) // 256 * 256
        self.o_ray = (self.o_t + w * self.cap * 4 + 255) // 256 * 256
        self.o_ready = (self.o_ray + w * self.cap * 4 + 255) // 256 * 256
        self.nbytes = (self.o_ready + w * 8 + 255) // 256 * 256
        self.wire = False
        lib = ctx._lib
        ptr, handle = C.c_void_p(), nat.IpcHandle()
        with torch.cuda.device(ctx.device):
            nat.check(ctx._h, lib.lrc_peer_buffer_create(ctx._h, self.nbytes, C.byref(ptr), C.byref(handle)))
        self.local_ptr = ptr.value
        handles = [None] * w
        if w > 1:
            dist.all_gather_object(handles, bytes(handle.bytes), group=group)
        else:
            handles[0] = bytes(handle.bytes)
        self.ptrs, self._opened = [], []
        for r in range(w):
            if r == self.rank:
                self.ptrs.append(self.local_ptr)
                continue
            h = nat.IpcHandle()
            C.memmove(C.byref(h), handles[r], 64)
            p = C.c_void_p()
            with torch.cuda.device(ctx.device):
                nat.check(ctx._h, lib.lrc_peer_buffer_open(ctx._h, C.byref(h), C.byref(p)))
            self.ptrs.append(p.value)
            self._opened.append(p.value)
        g = nat.Gather()
        g.n_targets = w
        for r in range(w):
            g.xyz[r] = self.ptrs[r]
            g.label[r] = self.ptrs[r] + self.o_lab
            g.frame_offset[r] = self.ptrs[r] + self.o_off
        g.point_base = self.rank * self.cap
        g.frame_base = self.rank * (self.pmax + 1)
        g.capacity = self.cap
        g.frame_capacity = self.pmax + 1      # a scan of more frames than this rank's offset region holds is refused
        self._g = g
        self.buffer = torch.as_tensor(_RawCudaBuffer(self.local_ptr, self.nbytes), device=ctx.device)
        if w > 1:
            dist.barrier(group=group)

    def local_out(self, frames: Optional[int] = None, incident: Optional[bool] = None) -> Dict[str, torch.Tensor]:
        """Output buffers for ``Context.scan_enqueue`` that ARE this rank's region of its own gather buffer: the compaction
        then writes the local cloud straight to its final place and the exchange kernel skips the copy to itself (one
        target and one pass over the local cloud less per step).  ``off`` is a private (P+1,) array -- the gathered
        offsets carry each rank's point base and are written by the exchange kernel."""
        xyz, lab, _ = self.views()
        nf = self.pmax if frames is None else int(frames)
        want_inc = (not self.wire) if incident is None else bool(incident)      # the compact wire format produces no local angles
        if want_inc and getattr(self, "_local_inc", None) is None:
            self._local_inc = torch.empty(self.cap, dtype=torch.float64, device=self.ctx.device)
        return {"xyz": xyz[self.rank], "incident": self._local_inc if want_inc else None, "prim": None, "label": lab[self.rank], "ray": None,
                "off": torch.zeros(nf + 1, dtype=torch.int64, device=self.ctx.device)}

    def enable(self, wire: bool = False, poses_all=None, frames_per_rank: Optional[List[int]] = None):
        """Switch the exchange on for the scans that follow.  ``wire=True`` selects the compact wire format
        (``lrc_set_gather_wire``): t | label | ray index travel (12 B per point instead of 16) and every rank rebuilds the
        other ranks' points on arrival from ``poses_all`` ((P_total,4,4), rank-major, rank r owning ``frames_per_rank[r]``
        of them) -- bit-identical, 25 % less NVLink traffic.  Collective when ``wire`` is set: every rank must call it, then
        issue the same sequence of scans, each of exactly its ``frames_per_rank`` frames, without an ``incident`` output
        (``local_out()`` leaves it out; ``incident()`` recomputes the angles of the whole cloud)."""
        import ctypes as C
        from . import _native as nat
        ctx = self.ctx
        nat.check(ctx._h, ctx._lib.lrc_set_gather(ctx._h, C.byref(self._g)))
        self.wire = bool(wire)
        if not wire:
            return
        w = self.world
        nfs = [self.pmax] * w if frames_per_rank is None else [int(x) for x in frames_per_rank]
        poses = np.ascontiguousarray(poses_all, dtype=np.float64).reshape(-1, 16)
        if len(poses) != sum(nfs) or max(nfs) > self.pmax:
            raise ValueError("poses_all must hold frames_per_rank[r] <= frames_per_rank of the constructor poses for every rank")
        with torch.cuda.device(ctx.device):
            self._wire_poses = torch.from_numpy(poses).to(ctx.device)
            # progress words of earlier scans must not satisfy the waits of the new ones
            torch.cuda.synchronize()
            if w > 1:
                dist.barrier(group=self.group)
            self.buffer[self.o_ready: self.o_ready + w * 8].zero_()
            torch.cuda.synchronize()
            if w > 1:
                dist.barrier(group=self.group)
        gw = nat.GatherWire()
        gw.enabled, gw.self_index = 1, self.rank
        p0 = 0
        for r in range(w):
            gw.t[r] = self.ptrs[r] + self.o_t
            gw.ray_idx[r] = self.ptrs[r] + self.o_ray
            gw.ready[r] = self.ptrs[r] + self.o_ready
            gw.rank_pose0[r], gw.rank_frames[r] = p0, nfs[r]
            gw.rank_point_base[r], gw.rank_frame_base[r] = r * self.cap, r * (self.pmax + 1)
            p0 += nfs[r]
        gw.all_poses = self._wire_poses.data_ptr()
        self._gw = gw
        nat.check(ctx._h, ctx._lib.lrc_set_gather_wire(ctx._h, C.byref(gw)))

    def disable(self):
        from . import _native as nat
        nat.check(self.ctx._h, self.ctx._lib.lrc_set_gather(self.ctx._h, None))
        self.wire = False

    def synchronize(self):
        """Make every rank's stores visible here: local stream sync, then a barrier across ranks."""
        torch.cuda.synchronize(self.ctx.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.wire and self.ctx.stat("wire_error"):
            raise RuntimeError("PeerGather: a peer's progress word did not arrive in time -- the ranks did not issue the same scans")

    def views(self):
        w = self.world
        xyz = self.buffer[: w * self.cap * 12].view(torch.float32).view(w, self.cap, 3)
        lab = self.buffer[self.o_lab: self.o_lab + w * self.cap * 4].view(torch.int32).view(w, self.cap)
        off = self.buffer[self.o_off: self.o_off + w * (self.pmax + 1) * 8].view(torch.int64).view(w, self.pmax + 1)
        return xyz, lab, off

    def incident(self, poses_all, frames_per_rank: Optional[List[int]] = None) -> torch.Tensor:
        """Incident angles of the WHOLE gathered cloud, recomputed on arrival: (world, cap) float64, row r holds the
        angles of rank r's points (valid up to that rank's point count).

        The reference's frame is ``(points, incident_angles)`` (raycast_engine_cpu.py:111); the angle is a pure function
        of the float32 point and the frame's sensor position (:95-107), so the exchange moves 16 B per point (xyz + label)
        and ``lrc_incident_angles`` redoes the float64 arithmetic here with the operations of the scan's own epilogue --
        the result is bit-identical to the angles the producing rank computed (``tests/test_gpu_multi.py``).
        ``poses_all``: the (P_total,4,4) poses in rank order, rank r owning ``frames_per_rank[r]`` of them (default:
        ``frames_per_rank`` of the constructor for every rank).  Call after ``synchronize()``."""
        import ctypes as C
        from . import _native as nat
        w = self.world
        nfs = [self.pmax] * w if frames_per_rank is None else [int(x) for x in frames_per_rank]
        poses = np.ascontiguousarray(poses_all, dtype=np.float64).reshape(-1, 16)
        if len(poses) != sum(nfs):
            raise ValueError("poses_all must hold one pose per gathered frame")
        ctx = self.ctx
        with torch.cuda.device(ctx.device):
            poses_d = torch.from_numpy(poses).to(ctx.device)
            if getattr(self, "_incident", None) is None:
                self._incident = torch.zeros((w, self.cap), dtype=torch.float64, device=ctx.device)
            xyz, _, off = self.views()
            p0 = 0
            for r in range(w):
                if nfs[r] > 0:
                    nat.check(ctx._h, ctx._lib.lrc_incident_angles(
                        ctx._h, C.c_void_p(xyz[r].data_ptr()), C.c_void_p(off[r].data_ptr()), nfs[r],
                        C.c_void_p(poses_d[p0:p0 + nfs[r]].data_ptr()), self.cap, C.c_void_p(self._incident[r].data_ptr()),
                        ctx._stream()))
                p0 += nfs[r]
            torch.cuda.current_stream(ctx.device).synchronize()       # poses_d may be freed after return
        return self._incident

    def assemble_numpy(self, frames_per_rank: Optional[List[int]] = None, poses_all=None) -> Dict[str, np.ndarray]:
        """Dense, pose-ordered cloud of all ranks (testing / export).  With ``poses_all`` the record is the reference's
        complete frame: ``incident`` is recomputed from (point, pose) on this rank (see ``incident``)."""
        xyz, lab, off = self.views()
        inc = None if poses_all is None else self.incident(poses_all, frames_per_rank)
        pts, labs, counts, incs = [], [], [], []
        for r in range(self.world):
            nf = self.pmax if frames_per_rank is None else frames_per_rank[r]
            o = off[r, : nf + 1].cpu().numpy() - r * self.cap
            m = int(o[-1])
            pts.append(xyz[r, :m].cpu().numpy()); labs.append(lab[r, :m].cpu().numpy().view(np.uint32))
            if inc is not None:
                incs.append(inc[r, :m].cpu().numpy())
            counts.append(np.diff(o))
        cnt = np.concatenate(counts)
        out = {"points": np.concatenate(pts), "label": np.concatenate(labs),
               "frame_offset": np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)}
        if inc is not None:
            out["incident"] = np.concatenate(incs)
        return out

    def close(self):
        lib = self.ctx._lib
        try:
            self.disable()
            with torch.cuda.device(self.ctx.device):
                torch.cuda.synchronize()
                if self.world > 1:
                    dist.barrier(group=self.group)
                for p in self._opened:
                    lib.lrc_peer_buffer_close(self.ctx._h, p)
                self._opened = []
                if self.world > 1:
                    dist.barrier(group=self.group)
                if self.local_ptr:
                    lib.lrc_peer_buffer_destroy(self.ctx._h, self.local_ptr)
                    self.local_ptr = None
        except Exception:
            pass
