#!/usr/bin/env python
"""
bench.py -- the reference's headline metric (Mrays/s and LiDAR frames/s of the ray-casting hot path) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c1|c4]

A "step" is one pass of the hot path over one batch: the whole trajectory of the workload (C2: 100 poses x
128000 rays of the dense 32-line sensor on the ~1M-triangle synthetic office, BASELINE.json configs[1]) through
ray generation -> LBVH traversal -> fused epilogue (labels, range filter, incident angle, ordered compaction).

  value     whole-job Mrays/s with mesh, BVH and poses resident in HBM (CUDA events around the K steps, max over ranks)
  e2e       same metric through the reference-facing API with HOST buffers: per step the mesh is uploaded and its
            LBVH rebuilt (the reference rebuilds its scene for every frame, raycast_engine_cpu.py:46-47; we do it once
            per trajectory), poses go H2D from pinned memory, points + incident angles + labels come back D2H
  e2e_per_frame   the reference's own call pattern (one engine.lidar_intersect_mesh(lidar, mesh) per waypoint,
            s3dis_simulator.py:254-263; numpy in, fresh numpy arrays out; mesh pinned with engine.set_mesh)
  roofline  the dominant kernel (k_trace) alone, against the resource that bounds it.  ncu shows issue slots and the
            L1 -> register return path, not HBM (profiles/): `achieved` = warp instructions of one launch (committed ncu
            capture of this workload and kernel build, profiles/traffic.json) / the kernel's device time measured live
            with CUDA events inside the library; `peak` = SMs x 4 schedulers x SM clock.  The sub-object `hbm` keeps
            SURVEY section 8d's algorithmic figure (node / triangle records fetched x 64 / 48 B + 24 B scratch per ray)
            next to the real DRAM bytes of one launch (`traffic`) and says what fraction of the measured HBM copy
            peak each is; `ncu` quotes the capture (issue_active, l1tex, dram, lanes per instruction).
  cpu_baseline  the CPU oracle (a scalar BVH2 port: Open3D/Embree is absent) on this box's cores, bounded sample
  extra     further BASELINE configs in the same run: C3 (BLK2GO, noise + labels, 256 poses in total, STRONG scaling over
            the N GPUs) at every N, C4 (500 poses on the 5M-triangle floor) at N = 8 (or with --extra c4)

N > 1 (torchrun): poses are sharded contiguously across ranks with a replicated mesh/BVH (weak scaling: 100
poses per GPU); every rank's compacted cloud is all-gathered inside the timed region -- by default by the library's
exchange kernel over NVLink peer memory (CUDA IPC), overlapped with traversal; --gather nccl uses NCCL instead.
After the timed steps every rank compares a pose subset of its gathered cloud (xyz, labels, offsets and the incident
angles recomputed on arrival) with a single-GPU scan of the same global poses: "gather_bit_identical".
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (mesh factory, target tris, sensor factory, poses per GPU, noise, description)
    "c1": dict(mesh="box_room", tris=50_000, sensor="8line", poses=1, noise=False,
               desc="single 8-line frame, ~50k-tri box room (BASELINE configs[0])"),
    "c2": dict(mesh="office", tris=1_000_000, sensor="32line", poses=100, noise=False,
               desc="32-line LiDAR, 100-waypoint trajectory, ~1M-tri synthetic office (BASELINE configs[1])"),
    "c3": dict(mesh="office", tris=1_000_000, sensor="blk2go", poses=256, noise=True,
               desc="dual-axis BLK2GO, 256 poses, angle noise + dropout + labels, ~1M tris (BASELINE configs[2])"),
    "c4": dict(mesh="floor_plan", tris=5_000_000, sensor="blk2go", poses=500, noise=True,
               desc="dual-axis BLK2GO, 500 poses, ~5M-tri multi-room floor (BASELINE configs[3])"),
}


def make_workload(lrc, name: str, world: int, tris_override=None, poses_override=None, mesh=None):
    """-> (workload dict, mesh, poses (P_total,4,4), intrinsics); P_total = poses per GPU x world (weak scaling)."""
    w = dict(WORKLOADS[name])
    if tris_override:
        w["tris"] = tris_override
    if poses_override:
        w["poses"] = poses_override
    syn = lrc.synthetic
    if mesh is None:
        mesh = getattr(syn, w["mesh"])(target_tris=w["tris"], seed=0)
    total_poses = w["poses"] * world
    if w["mesh"] == "office":
        wps = syn.office_waypoints(total_poses)
    elif w["mesh"] == "floor_plan":
        wps = syn.floor_plan_waypoints(total_poses)
    else:
        wps = [lrc.Waypoint(3.137 + 0.01 * k, 2.718, 1.0, 0.3) for k in range(total_poses)]
    poses = lrc.poses_from_waypoints(wps)
    intr = {"8line": lrc.Indoor8LineLidarIntrinsics.create_standard_8line,
            "32line": lrc.Indoor8LineLidarIntrinsics.create_dense_32line,
            "blk2go": lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis}[w["sensor"]]()
    return w, mesh, poses, intr


def workload_config(args, w, n_tris: int, n_frame: int, world: int) -> dict:
    """The `config` object: identical for the two arms (--impl ours / reference) given the same command line."""
    return {"workload": f"{args.workload}: {w['desc']}", "tris": int(n_tris), "rays_per_frame": int(n_frame),
            "poses_per_gpu": int(w["poses"]), "poses_total": int(w["poses"] * world), "noise": bool(w["noise"]),
            "parallelism": f"pose-sharded x{world}, replicated BVH",
            "l2": "flushed between timed iterations (256 MiB fill); inputs resident in HBM for `value`, host buffers for `e2e`"}


# ---- clocks sampler ----------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index: int, period: float = 0.005):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            self._halt.wait(self.period)

    def finish(self) -> dict:
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def physical_gpu_index(local_index: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:
            return local_index
    return local_index


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def host_threads() -> int:
    """Cores this process may use.  torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must not inherit that."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


CPU_ARM_NOTE = ("oracle C port stands in for Open3D/Embree (absent, not installable): a scalar BVH2 traverser with a task-parallel "
                "binned-SAH build -- Embree is typically several times faster per core, so this ratio is not a statement about Open3D; "
                "the ray table comes from the oracle's C generator, not from the reference's Python get_rays (113 ms per 32-line frame, "
                "BASELINE.md), because /root/reference does not exist on the GPU box")


# ---- CPU side: the oracle as stand-in for the reference's CPU engine -----------------------------------
def cpu_frame_reference_structured(orc, mesh, pose, intr, lrc):
    """One frame exactly as reference raycast_engine_cpu.py:75-111 does it: rays, NEW scene, cast, numpy epilogue."""
    class _L:
        pass
    lid = _L()
    lid.pose, lid.intrinsics = pose, intr
    if hasattr(intr, "num_vertical_lines"):
        lid.get_rays = lambda: orc.gen_rays_dual_axis(pose, orc.dual_params(intr), seed=2, pose_idx=0)[0]
    else:
        lid.get_rays = lambda: orc.gen_rays_single_axis(pose, intr.vertical_degrees, intr.horizontal_res)
    pts, inc = orc.OracleEngineCPU().lidar_intersect_mesh(lid, (mesh.vertices, mesh.triangles))
    return len(pts)


def cpu_baseline(lrc, mesh, poses, intr, n_rays_frame, budget_s=20.0):
    from oracle import oracle as orc
    orc.set_num_threads(host_threads())
    cores = orc.num_threads()
    # (i) reference-structured: scene rebuilt per frame
    t0 = time.perf_counter()
    frames = 0
    while frames < 2 or (time.perf_counter() - t0 < budget_s * 0.5 and frames < len(poses)):
        cpu_frame_reference_structured(orc, mesh, poses[frames % len(poses)], intr, lrc)
        frames += 1
    dt = time.perf_counter() - t0
    structured = frames * n_rays_frame / dt / 1e6
    # (ii) cast only: scene prebuilt, C epilogue, all cores
    scene = orc.OracleScene((mesh.vertices, mesh.triangles))
    dual = hasattr(intr, "num_vertical_lines")
    t0 = time.perf_counter()
    f2, rays_done = 0, 0
    while f2 < 3 or (time.perf_counter() - t0 < budget_s * 0.5 and f2 < len(poses)):
        pose = poses[f2 % len(poses)]
        if dual:
            rays, keep = orc.gen_rays_dual_axis(pose, orc.dual_params(intr), seed=2, pose_idx=f2, compact=True)
        else:
            rays = orc.gen_rays_single_axis(pose, intr.vertical_degrees, intr.horizontal_res)
        t, pid = scene.cast_rays(rays)
        orc.epilogue_c(rays, t, pid, center=pose[:3, 3], max_range=intr.max_range, tri_label=mesh.triangle_labels)
        rays_done += len(rays)
        f2 += 1
    dt2 = time.perf_counter() - t0
    stats = scene.stats()
    return {
        "value": round(structured, 4), "unit": "Mrays/s", "cores": cores, "kind": "port",
        "sample": f"{frames} frames of the trajectory, reference-structured (ray table + scene rebuild + cast + numpy epilogue per frame, "
                  f"as raycast_engine_cpu.py:46-47 does); " + CPU_ARM_NOTE,
        "frames_per_s": round(frames / dt, 4),
        "cast_only": {"value": round(rays_done / dt2 / 1e6, 4), "unit": "Mrays/s", "frames": f2,
                      "note": "scene prebuilt once, C epilogue, all cores",
                      "box_tests_per_ray": round(stats["box_tests"] / max(1, stats["rays"]), 2),
                      "tri_tests_per_ray": round(stats["tri_tests"] / max(1, stats["rays"]), 2)},
    }


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    world = max(1, args.gpus)
    import lrc_b200 as lrc
    from oracle import oracle as orc
    orc.set_num_threads(host_threads())            # torchrun pins OMP_NUM_THREADS=1; this arm uses every core it may
    cores = orc.num_threads()
    w, mesh, poses, intr = make_workload(lrc, args.workload, world, args.tris, args.poses)
    n_frame = lrc.rays_per_frame(intr)
    k = 0
    t_cal = time.perf_counter()
    for _ in range(max(1, args.warmup)):
        cpu_frame_reference_structured(orc, mesh, poses[k % len(poses)], intr, lrc)
        k += 1
    t_frame = (time.perf_counter() - t_cal) / max(1, args.warmup)
    # a step = a bounded sample of the workload: as many frames of the trajectory as fit the budget, cycling through ALL poses
    if args.ref_frames > 0:
        frames_per_step = args.ref_frames
    else:
        frames_per_step = int(max(1, min(len(poses), args.ref_budget / max(1e-6, t_frame * max(1, args.steps)))))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for _ in range(frames_per_step):
            cpu_frame_reference_structured(orc, mesh, poses[k % len(poses)], intr, lrc)
            k += 1
    dt = time.perf_counter() - t0
    rays = args.steps * frames_per_step * n_frame
    val = rays / dt / 1e6
    passes = args.steps * frames_per_step / len(poses)
    sample = (f"{frames_per_step} frame(s) per step, consecutive poses cycling through all {len(poses)} poses of the trajectory "
              f"({passes:.2f} passes in the timed region); every frame = ray table + scene rebuild + cast + numpy epilogue "
              f"(reference raycast_engine_cpu.py:75-111); {cores} threads; " + CPU_ARM_NOTE)
    line = {
        "impl": "reference", "metric": "lidar_raycast_throughput", "value": round(val, 4), "unit": "Mrays/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, w, len(mesh.triangles), n_frame, world),
        "frames_per_s": round(args.steps * frames_per_step / dt, 4),
        "cpu_baseline": {"value": round(val, 4), "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample,
                         "frames_per_step": frames_per_step},
        "e2e": {"value": round(val, 4), "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---- our arm -------------------------------------------------------------------------------------------
def build_tag(ctx) -> str:
    """Names the kernel build / tree options a profile was taken with; bench.py only quotes a capture of the same tag."""
    return "fmt%d-q%d-leaf%d-var%d-tune%d-wp%d-rpt%d-pers%d" % tuple(
        ctx.stat(k) for k in ("node_format", "build_quality", "leaf_size", "variant", "tune", "warp_packet", "rays_per_thread", "persistent"))


def load_capture(workload: str, n_tris: int, rays_per_launch: int, tag: str):
    """The committed ncu capture of k_trace for this workload (profiles/traffic.json), or (None, why)."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tpath):
        return None, "profiles/traffic.json missing"
    try:
        tj = json.load(open(tpath)).get(workload)
    except Exception as e:
        return None, f"profiles/traffic.json unreadable: {e}"
    if not tj:
        return None, f"no capture for workload {workload}"
    if int(tj.get("tris", 0)) != int(n_tris):
        return None, "capture is of another mesh size"
    if tj.get("build_tag") != tag:
        return None, f"capture is of another kernel build ({tj.get('build_tag')} vs {tag})"
    if int(tj.get("rays_per_launch", 0)) != int(rays_per_launch):
        # same mesh, sensor and kernel, another number of frames per launch (pose chunks at N > 1): the per-launch totals of
        # the capture are scaled by the ray count -- instructions and bytes per ray do not depend on how many frames a launch holds
        f = float(rays_per_launch) / float(tj["rays_per_launch"])
        tj = dict(tj)
        for k in ("inst_executed", "dram_bytes_read", "dram_bytes_write", "dram_bytes_per_launch", "lts_t_bytes", "l1tex_t_bytes"):
            if k in tj:
                tj[k] = int(tj[k] * f)
        tj.pop("kernel_ms_ncu", None)
        tj["source"] = f"scaled x{f:.4f} by rays per launch from: " + str(tj.get("source"))
    return tj, None


class Leg:
    """One workload on this rank: resident inputs, timed device steps, kernel split."""

    def __init__(self, lrc, torch, ctx, dev, name, w, mesh, poses_all, intr, rank, world, strong: bool):
        self.lrc, self.torch, self.ctx, self.dev = lrc, torch, ctx, dev
        self.name, self.w, self.mesh, self.poses_all, self.intr = name, w, mesh, poses_all, intr
        self.rank, self.world = rank, world
        self.n_frame = lrc.rays_per_frame(intr)
        self.shard = lrc.shard_range(len(poses_all), rank, world)
        self.poses = poses_all[self.shard.start:self.shard.stop]
        self.P = len(self.poses)
        self.pmax = max(len(lrc.shard_range(len(poses_all), r, world)) for r in range(world))
        self.noise = lrc.NoiseConfig.from_intrinsics(intr, seed=2, pose_index_base=self.shard.start) if w["noise"] else None
        self.poses_d = torch.from_numpy(np.ascontiguousarray(self.poses.reshape(-1, 16))).to(dev)
        self.bufs, _ = ctx._alloc_out(max(1, self.P * self.n_frame), self.P)
        self.peer = None
        self.wire = False

    def enable_gather(self, args):
        from lrc_b200.distributed import PeerGather
        self.peer = PeerGather(self.ctx, cap_per_rank=self.pmax * self.n_frame, frames_per_rank=self.pmax)
        # a chunk should stay a launch of at least ~1M rays (a few waves of blocks): small shards get fewer chunks
        self.chunks = int(max(1, min(args.gather_chunks, (self.pmax * self.n_frame) // 1_000_000)))
        self.ctx.set_option("gather_chunks", self.chunks)
        if args.push_blocks:
            self.ctx.set_option("push_blocks", args.push_blocks)
        if args.gather_ramp:
            self.ctx.set_option("gather_ramp", args.gather_ramp)
        if args.gather_taper:
            self.ctx.set_option("gather_taper", args.gather_taper)
        if args.push_mode is not None:
            self.ctx.set_option("push_mode", args.push_mode)
        if args.push_tile:
            self.ctx.set_option("push_tile", args.push_tile)
        # compact wire format (t | label | ray index, 12 B per point; points rebuilt on arrival), "--wire 1".  Measured equal to
        # xyz | label at 8 GPUs (3.05 vs 3.07 ms per step: what the link saves the rebuild kernels spend) and slower below, so off
        # by default (profiles/r02h_scaling.md)
        self.wire = bool(args.wire == 1)
        if self.wire:
            nfs = [len(self.lrc.shard_range(len(self.poses_all), r, self.world)) for r in range(self.world)]
            self.peer.enable(wire=True, poses_all=self.poses_all, frames_per_rank=nfs)
            self.bufs = self.peer.local_out(self.P)
        else:
            self.peer.enable()
            if not args.gather_copy_self:
                self.bufs = self.peer.local_out(self.P)    # compact straight into this rank's region of its own gather buffer

    def step(self):
        self.ctx.scan_enqueue(self.poses_d, self.intr, self.noise, self.bufs)

    def counters(self):
        ctx = self.ctx
        ctx.set_counting(True)
        ctx.counters(reset=True)
        self.step()
        cnt = ctx.counters(reset=True)
        ctx.set_counting(False)
        self.total_pts = int(self.bufs["off"][-1].item())
        rays = max(1, cnt["rays"])
        return {"rays_cast": cnt["rays"], "nodes_per_ray": cnt["nodes_visited"] / rays, "tris_per_ray": cnt["tris_tested"] / rays,
                "hit_fraction": self.total_pts / rays}

    def timed(self, steps, warmup, flush, dist):
        """-> (ms per step, max over ranks), launches, wall seconds"""
        torch = self.torch
        for k in range(max(warmup, 3)):
            flush.fill_(k & 255)
            self.step()
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        launches0 = self.ctx.launch_count()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        wall0 = time.perf_counter()
        for s in range(steps):
            flush.fill_(s & 255)                               # evict the BVH from L2 between timed iterations
            starts[s].record()
            self.step()
            ends[s].record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
        if self.world > 1:
            dist.barrier()
        launches = self.ctx.launch_count() - launches0
        dev_ms = sum(starts[s].elapsed_time(ends[s]) for s in range(steps))
        t = torch.tensor([dev_ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / steps, launches, wall

    def kernel_split(self, reps, flush):
        """Device time of k_trace and of compaction (+ exchange) per launch, CUDA events inside the library."""
        torch, ctx = self.torch, self.ctx
        torch.cuda.synchronize()
        ctx.set_option("kernel_timing", 1)
        tr, cp, st, n_launch = [], [], [], 1
        for s in range(reps):
            flush.fill_(s & 255)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            self.step()
            b.record()
            torch.cuda.synchronize()
            kt = ctx.kernel_times()
            n_launch = max(1, kt["trace_launches"])
            tr.append(kt["trace_ms"]); cp.append(kt["compact_ms"]); st.append(a.elapsed_time(b))
        ctx.set_option("kernel_timing", 0)
        return {"trace_ms_per_step": float(np.mean(tr)), "after_trace_ms_per_step": float(np.mean(cp)), "launches_per_step": int(n_launch),
                "step_ms_same_runs": float(np.mean(st))}

    def verify_gather(self, dist) -> dict:
        """Every rank: a pose subset of the cloud it GATHERED (all ranks' frames) against a local single-GPU scan of the
        same global poses -- xyz, labels, frame sizes and the incident angles recomputed on arrival, bit for bit."""
        lrc, torch = self.lrc, self.torch
        peer = self.peer
        peer.synchronize()
        peer.disable()
        nfs = [len(lrc.shard_range(len(self.poses_all), r, self.world)) for r in range(self.world)]
        xyz, lab, off = peer.views()
        inc = peer.incident(self.poses_all, nfs)
        off_h = off.cpu().numpy()
        ok, checked, pts_checked = True, 0, 0
        for r in range(self.world):
            g0 = lrc.shard_range(len(self.poses_all), r, self.world).start
            for f in sorted({0, nfs[r] // 2, nfs[r] - 1}):
                if f < 0 or f >= nfs[r]:
                    continue
                nz = None
                if self.noise is not None:
                    nz = lrc.NoiseConfig(self.noise.angle_noise_std, self.noise.dropout_probability, self.noise.range_noise_std,
                                         self.noise.seed, g0 + f)
                ref = self.ctx.scan(self.poses_all[g0 + f][None], self.intr, nz)
                a, b = int(off_h[r, f]) - r * peer.cap, int(off_h[r, f + 1]) - r * peer.cap
                same = (b - a) == ref.num_points
                if same and b > a:
                    same = (torch.equal(xyz[r, a:b], ref.points) and torch.equal(lab[r, a:b], ref.label)
                            and torch.equal(inc[r, a:b], ref.incident))
                ok = ok and bool(same)
                checked += 1
                pts_checked += max(0, b - a)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.dev)
        if self.world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        return {"ok": bool(flag.item()), "frames_checked_per_rank": checked, "points_checked_rank0": pts_checked}

    def close(self):
        if self.peer is not None:
            self.peer.close()
            self.peer = None


def extra_leg(lrc, torch, dist, ctx, dev, args, name, rank, world, flush, mesh_cache):
    """A further BASELINE config inside the same run: FIXED total trajectory split over the ranks (strong scaling)."""
    w = dict(WORKLOADS[name])
    key = (w["mesh"], w["tris"])
    if key not in mesh_cache:
        mesh_cache.clear()
        mesh_cache[key] = getattr(lrc.synthetic, w["mesh"])(target_tris=w["tris"], seed=0)
        v, f, lab = lrc.mesh_arrays(mesh_cache[key])
        ctx.set_mesh_arrays(v, f, lab)
    mesh = mesh_cache[key]
    wps = lrc.synthetic.office_waypoints(w["poses"]) if w["mesh"] == "office" else lrc.synthetic.floor_plan_waypoints(w["poses"])
    poses_all = lrc.poses_from_waypoints(wps)
    intr = lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()
    leg = Leg(lrc, torch, ctx, dev, name, w, mesh, poses_all, intr, rank, world, strong=True)
    if world > 1 and args.gather == "p2p":
        leg.enable_gather(args)
    cnt = leg.counters()
    steps = max(3, min(args.steps, 10))
    ms, launches, _ = leg.timed(steps, 3, flush, dist)
    split = leg.kernel_split(3, flush)
    out = {"workload": f"{name}: {w['desc']}", "scaling": "strong", "poses_total": len(poses_all), "poses_this_rank": leg.P,
           "tris": int(len(mesh.triangles)), "rays_per_frame": leg.n_frame, "steps": steps,
           "value": round(leg.n_frame * len(poses_all) / (ms * 1e-3) / 1e6, 2), "unit": "Mrays/s", "ms_per_step": round(ms, 4),
           "frames_per_s": round(len(poses_all) / (ms * 1e-3), 1),
           "k_trace_ms_rank0": round(split["trace_ms_per_step"], 4),
           ("compact_plus_exchange_ms_rank0" if leg.peer is not None else "compact_ms_rank0"): round(split["after_trace_ms_per_step"], 4),
           "chunks_per_step": split["launches_per_step"], "nodes_per_ray": round(cnt["nodes_per_ray"], 2),
           "tris_per_ray": round(cnt["tris_per_ray"], 2), "points_per_step_rank0": leg.total_pts, "gpu_launches": int(launches)}
    if leg.peer is not None:
        out["gather_bit_identical"] = leg.verify_gather(dist)["ok"]
    leg.close()
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    import lrc_b200 as lrc

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1 and not args.no_numa_bind:
        from lrc_b200.distributed import bind_to_gpu_numa_node
        numa = bind_to_gpu_numa_node(local)      # before any pinned allocation: staging buffers on the GPU's own socket
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w, mesh, poses_all, intr = make_workload(lrc, args.workload, world, args.tris, args.poses)
    engine = lrc.RaycastEngineGPU(device=local)
    ctx = engine.ctx
    for key, val in (("variant", args.variant), ("node_format", args.node_format), ("l2_persist", args.l2_persist), ("tune", args.tune)):
        if val is not None:
            ctx.set_option(key, val)
    verts, tris, labels = lrc.mesh_arrays(mesh)

    # ---- one-off: BVH build time (CUDA events), resident inputs ----
    v_d = torch.from_numpy(verts).to(dev)
    f_d = torch.from_numpy(tris).to(dev)
    l_d = torch.from_numpy(labels.view(np.int32)).to(dev)
    ctx.set_mesh_arrays(v_d, f_d, l_d)                     # warm-up build (allocations)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ctx.set_mesh_arrays(v_d, f_d, l_d)
    e1.record()
    torch.cuda.synchronize()
    bvh_ms = e0.elapsed_time(e1)
    info = ctx.bvh_info()
    tag = build_tag(ctx)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2

    leg = Leg(lrc, torch, ctx, dev, args.workload, w, mesh, poses_all, intr, rank, world, strong=False)
    n_frame, P = leg.n_frame, leg.P
    cnt = leg.counters()                                   # untimed, counting instantiation of the traversal kernel
    sharded = None
    if world > 1 and args.gather == "nccl":
        from lrc_b200.distributed import OverlappedShardedScan
        sharded = OverlappedShardedScan(ctx, leg.poses, intr, leg.noise, chunks=args.gather_chunks)
        leg.step = sharded.step     # chunked scan; each chunk's xyz|label|offsets block is all-gathered (NCCL, async) behind the next chunk
    elif world > 1:
        # fused compaction + all-gather: an exchange kernel stores every compacted chunk into all ranks' buffers over
        # NVLink peer memory while the next chunk is traversed (lrc_set_gather)
        leg.enable_gather(args)

    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    ms_per_step, launches, wall = leg.timed(args.steps, args.warmup, flush, dist)
    clocks = sampler.finish()
    rays_per_step_all = n_frame * len(poses_all)           # dense rays generated per step, all ranks
    value = rays_per_step_all / (ms_per_step * 1e-3) / 1e6

    # ---- the dominant kernel alone: CUDA events recorded inside the library on the stream k_trace is launched on ----
    split = leg.kernel_split(max(3, min(args.steps, 10)), flush)
    n_launch = split["launches_per_step"]
    trace_ms = split["trace_ms_per_step"] / n_launch       # average duration of ONE k_trace launch
    rays_per_launch = cnt["rays_cast"] / n_launch
    peak_hbm, peak_src = measured_peak_gbs()
    # SURVEY section 8d's algorithmic bytes of one launch: node + triangle records fetched, plus the 24 B of scratch it writes per ray
    trace_bytes_per_ray = cnt["nodes_per_ray"] * 64 + cnt["tris_per_ray"] * 48 + 16 + 8
    alg_bytes = rays_per_launch * trace_bytes_per_ray
    alg_gbs = alg_bytes / (trace_ms * 1e-3) / 1e9
    cap, cap_why = load_capture(args.workload, len(tris), int(rays_per_launch), tag)
    num_sms = torch.cuda.get_device_properties(dev).multi_processor_count
    sm_mhz = clocks.get("sm_max_mhz") or 1965
    issue_peak = num_sms * 4 * sm_mhz * 1e6 / 1e9             # G warp-instructions / s
    footprint_mb = (info["bytes_nodes"] + info["bytes_tris"]) / 1e6
    roofline = {
        "bound": "issue", "kernel": "k_trace", "kernel_ms": round(trace_ms, 4), "launches_per_step": int(n_launch),
        "achieved": None, "peak": round(issue_peak, 1), "unit": "Gwarp-inst/s", "frac": None, "traffic": None,
        "peak_source": f"{num_sms} SMs x 4 warp schedulers x {sm_mhz} MHz (max SM clock; one warp instruction per scheduler and cycle)",
        "why_not_hbm": "rays of a warp are adjacent beams of one scan line, so node / triangle records are served by L1 and L2; ncu shows the "
                       "kernel co-limited by issue slots and the L1 -> register return path with DRAM under 10 % busy (see `ncu`); the "
                       "SURVEY 8d byte model is kept under `hbm`",
        "hbm": {"algorithmic_bytes_per_launch": int(alg_bytes), "bytes_per_ray": round(trace_bytes_per_ray, 1),
                "algorithmic_gbs": round(alg_gbs, 1), "hbm_peak_gbs": peak_hbm, "peak_source": peak_src,
                "algorithmic_over_hbm_peak": round(alg_gbs / peak_hbm, 4),
                "note": f"algorithmic bytes = every node (64 B) and triangle (48 B) record a ray fetches (COUNT instantiation of k_trace) + 24 B "
                        f"scratch written per ray; most fetches are L1/L2 hits (BVH {footprint_mb:.0f} MB vs 126 MB L2), so this ratio is not an HBM "
                        f"utilisation and may exceed 1; dram_gbs / dram_frac are the real DRAM bytes of the capture over the live kernel time"},
        "nodes_per_ray": round(cnt["nodes_per_ray"], 2), "tris_per_ray": round(cnt["tris_per_ray"], 2),
        "hit_fraction": round(cnt["hit_fraction"], 4), "build_tag": tag,
        ("compact_plus_exchange_ms" if leg.peer is not None else "compact_ms"): round(split["after_trace_ms_per_step"] / n_launch, 4),
        "step_ms_same_runs": round(split["step_ms_same_runs"], 4),
    }
    if cap is not None:
        ach = cap["inst_executed"] / (trace_ms * 1e-3) / 1e9
        roofline.update({"achieved": round(ach, 1), "frac": round(ach / issue_peak, 4), "traffic": int(cap["dram_bytes_per_launch"]),
                         "traffic_source": cap.get("source")})
        roofline["hbm"].update({"dram_gbs": round(cap["dram_bytes_per_launch"] / (trace_ms * 1e-3) / 1e9, 1),
                                "dram_frac_of_hbm_peak": round(cap["dram_bytes_per_launch"] / (trace_ms * 1e-3) / 1e9 / peak_hbm, 4)})
        roofline["ncu"] = {k: cap[k] for k in ("issue_active_pct", "l1tex_pct", "lsu_writeback_pct", "dram_pct", "lts_pct", "simt_lanes",
                                              "l1_hit_pct", "l2_hit_pct", "inst_executed", "kernel_ms_ncu", "dram_bytes_read",
                                              "dram_bytes_write", "lts_t_bytes", "l1tex_t_bytes") if k in cap}
    else:
        roofline["capture"] = cap_why

    exchange = None
    gather_check = None
    if leg.peer is not None:
        gather_check = leg.verify_gather(dist)
        pts = leg.total_pts
        bpp = 12 if leg.wire else 16
        exchange = {"bytes_out_per_gpu_per_step": int(pts * bpp * (world - 1)), "bytes_in_per_gpu_per_step": int(pts * bpp * (world - 1)),
                    "bytes_per_point_on_the_wire": bpp,
                    "record": ("t 4 B + label 4 B + ray index 4 B per point + frame offsets; every rank rebuilds the other ranks' xyz on arrival "
                               "(same ray generation and float32 point arithmetic as the scan, bit-identical) and recomputes incident angles on demand"
                               if leg.wire else
                               "xyz 12 B + label 4 B per point + frame offsets; incident angles are recomputed on arrival (lrc_incident_angles)"),
                    "nvlink_ingest_gbs_over_step": round(pts * bpp * (world - 1) / (ms_per_step * 1e-3) / 1e9, 1),
                    "nvlink_ingest_gbs_over_exchange_kernels": round(pts * bpp * (world - 1) / max(1e-9, split["after_trace_ms_per_step"] * 1e-3) / 1e9, 1)}

    # ---- end to end through the reference-facing API with host buffers ----
    pinned_pose = torch.from_numpy(np.ascontiguousarray(leg.poses.reshape(-1, 16))).pin_memory()
    pv, pf, pl = (torch.from_numpy(verts).pin_memory(), torch.from_numpy(tris).pin_memory(),
                  torch.from_numpy(labels.view(np.int32)).pin_memory())
    host = ctx.alloc_host_buffers(P * n_frame, P, labels=True)

    def e2e_step(c=ctx, hb=host):
        # mesh H2D + LBVH build (the reference rebuilds its scene on every frame; here once per trajectory) ...
        c.set_mesh_host(pv, pf, pl)
        # ... then the trajectory: poses H2D from pinned memory, chunked scan, D2H pipelined behind later chunks
        res = c.scan_to_host(pinned_pose, intr, leg.noise, host=hb, chunk_poses=args.e2e_chunk)
        return res["num_points"]

    for _ in range(2):
        m = e2e_step()
    if world > 1:
        dist.barrier()
    n_e2e = max(4, min(args.steps, 10)) // 2 * 2
    # (a) one call at a time: the latency of a trajectory call
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        m = e2e_step()
    lat_s = (time.perf_counter() - t0) / n_e2e
    # (b) the same calls from `--e2e-inflight` host threads, each with its own engine context and host buffers: trajectories
    # (rooms) are independent, so the mesh upload + LBVH build of one call overlaps the PCIe read-back of the other; every
    # step still uploads its mesh and poses and returns all of its points, angles and labels inside the timed region
    thr_s = lat_s
    if args.e2e_inflight > 1:
        ctxs = [ctx] + [lrc.Context(local) for _ in range(args.e2e_inflight - 1)]
        hosts = [host] + [ctx.alloc_host_buffers(P * n_frame, P, labels=True) for _ in range(args.e2e_inflight - 1)]
        for c in ctxs[1:]:
            for key, val in (("variant", args.variant), ("node_format", args.node_format), ("tune", args.tune)):
                if val is not None:
                    c.set_option(key, val)

        def worker(k, steps):
            torch.cuda.set_device(local)
            for _ in range(steps):
                e2e_step(ctxs[k], hosts[k])
        for k in range(1, len(ctxs)):
            worker(k, 2)                                       # allocations, first build
        per = n_e2e // len(ctxs)
        if world > 1:
            dist.barrier()
        threads = [threading.Thread(target=worker, args=(k, per)) for k in range(len(ctxs))]
        t0 = time.perf_counter()
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        torch.cuda.synchronize()
        thr_s = (time.perf_counter() - t0) / (per * len(ctxs))
        for c in ctxs[1:]:
            c.close()
        del hosts
    e2e_s = min(lat_s, thr_s)
    t = torch.tensor([e2e_s, lat_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s, lat_s = float(t[0].item()), float(t[1].item())
    h2d = pv.numel() * 4 + pf.numel() * 4 + pl.numel() * 4 + pinned_pose.numel() * 8
    d2h = m * (12 + 8 + 4) + 8
    e2e = {"value": round(rays_per_step_all / e2e_s / 1e6, 2), "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d),
           "d2h_bytes_per_step": int(d2h), "ms_per_step": round(e2e_s * 1e3, 3),
           "calls_in_flight": args.e2e_inflight if thr_s < lat_s else 1,
           "single_call": {"value": round(rays_per_step_all / lat_s / 1e6, 2), "ms_per_step": round(lat_s * 1e3, 3),
                           "note": "one trajectory call at a time (latency of a call)"},
           "includes": "per step: mesh upload + LBVH build + pose upload + scan + D2H of points/incident/labels; "
                       f"{args.e2e_inflight} host threads issue these calls on their own engine contexts, so one call's upload + build overlaps "
                       "the other's PCIe read-back"}
    # the same call with the incident angles left on the device (16 instead of 24 B per point over PCIe; the statistics that
    # consume them run on the GPU): reported beside the full record, never instead of it
    if rank == 0 and world == 1:
        host16 = ctx.alloc_host_buffers(P * n_frame, P, labels=True, incident=False)

        def e2e16_step():
            ctx.set_mesh_host(pv, pf, pl)
            return ctx.scan_to_host(pinned_pose, intr, leg.noise, host=host16, chunk_poses=args.e2e_chunk)["num_points"]
        for _ in range(2):
            e2e16_step()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            m16 = e2e16_step()
        s16 = (time.perf_counter() - t0) / n_e2e
        e2e["points_and_labels_only"] = {"value": round(rays_per_step_all / s16 / 1e6, 2), "unit": "Mrays/s", "ms_per_step": round(s16 * 1e3, 3),
                                         "d2h_bytes_per_step": int(m16 * 16 + 8),
                                         "note": "incident angles stay on the device (alloc_host_buffers(incident=False))"}
        del host16

    # ---- the reference's own call pattern: one lidar_intersect_mesh per waypoint (rank 0, N = 1) ----
    per_frame = None
    if rank == 0 and world == 1 and not args.no_per_frame:
        from tools.frame_latency import per_frame_latency
        pf_poses = poses_all[: min(len(poses_all), 45)] if len(poses_all) >= 12 else np.repeat(poses_all, 12, axis=0)
        r = per_frame_latency(engine, mesh, intr, pf_poses, pinned=True, warm=5)
        per_frame = {"value": round(r["Mrays_per_s"], 2), "unit": "Mrays/s", "ms_per_frame": round(r["ms_median"], 4),
                     "ms_per_frame_p90": round(r["ms_p90"], 4), "frames_per_s": round(r["frames_per_s"], 1), "frames_timed": r["frames"],
                     "call": "engine.set_mesh(mesh) once, then create_lidar(cfg, pose) + engine.lidar_intersect_mesh(lidar, mesh) per waypoint "
                             "(s3dis_simulator.py:254-263): pose from host memory, fresh numpy points + incident angles back, one synchronisation per frame",
                     "h2d_bytes_per_frame": 128, "d2h_bytes_per_frame": int(r["rays_per_frame"] * 20 + 16)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(lrc, mesh, leg.poses, intr, n_frame, budget_s=args.cpu_budget)

    # ---- further BASELINE configs in the same run ----
    extras = {}
    want = [x for x in (args.extra or "").split(",") if x]
    if not want and args.workload == "c2" and not args.tris and not args.poses:
        want = ["c3"] + (["c4"] if world >= 8 else [])
    leg.close()
    if want and want != ["none"]:
        mesh_cache = {(w["mesh"], w["tris"]): mesh}
        ctx.set_mesh_arrays(v_d, f_d, l_d)                 # the e2e leg re-uploaded the same mesh; make the state explicit
        for name in want:
            if name in WORKLOADS and name not in ("c1", "c2"):
                extras[name] = extra_leg(lrc, torch, dist, ctx, dev, args, name, rank, world, flush, mesh_cache)

    line = None
    leg_wire = leg.wire
    if rank == 0:
        cfg = workload_config(args, w, len(tris), n_frame, world)
        line = {
            "metric": "lidar_raycast_throughput", "value": round(value, 2), "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "run": {"numa_node_rank0": numa, "l2_persist_pct": int(args.l2_persist) if args.l2_persist is not None else ctx.default_l2_persist(),
                    "build_tag": tag, "host_threads": host_threads(),
                    "collective": ("none" if world == 1 else
                                   f"all-gather over NVLink peer memory: each compacted pose chunk's {'t|label|ray index' if leg_wire else 'xyz|label'}|frame_offset is pushed to the other {world - 1} ranks by an "
                                   f"exchange kernel (TMA bulk copies global -> shared -> peer; --push-mode 0: 16 B vector stores) while the next chunk is traversed"
                                   f"{'; the other ranks rebuild xyz on arrival (release/acquire progress words, own stream), still' if leg_wire else ';'} "
                                   f"{args.gather_chunks} chunks, inside the step" if args.gather == "p2p" else
                                   f"NCCL all-gather of xyz|label|frame_offset blocks, {args.gather_chunks} chunks, overlapped with traversal, inside the step"),
                    "process_group": "none" if world == 1 else "nccl (barrier, all_reduce of timings and of the gather check)"},
            "frames_per_s": round(len(poses_all) / (ms_per_step * 1e-3), 1),
            "wall_ms_per_step_incl_flush": round(wall / args.steps * 1e3, 4),
            "bvh_build_ms": round(bvh_ms, 3), "bvh": {k: info[k] for k in ("num_nodes", "max_depth", "sah_cost", "bytes_nodes", "bytes_tris")},
            "points_per_step_rank0": leg.total_pts,
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
        }
        if per_frame is not None:
            line["e2e_per_frame"] = per_frame
        if exchange is not None:
            line["exchange"] = exchange
        if gather_check is not None:
            line["gather_bit_identical"] = gather_check["ok"]
            line["gather_check"] = gather_check
        if extras:
            line["extra"] = extras
        if cpu is not None:
            line["cpu_baseline"] = cpu
    bad = (gather_check is not None and not gather_check["ok"]) or any(e.get("gather_bit_identical") is False for e in extras.values())
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        if world > 1:
            time.sleep(1.0)                                # the other ranks' NCCL log lines (NCCL_DEBUG=INFO) first ...
        print(json.dumps(line), flush=True)
        if world > 1:
            # ... and the JSON line last: NCCL writes a closing line from a library destructor at process exit, after
            # anything Python can print -- leave without running C-level exit handlers (everything is flushed and joined)
            sys.stdout.flush()
            sys.stderr.flush()
            os._exit(1 if bad else 0)
    return 1 if bad else 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--tris", type=int, default=None, help="override the triangle count (debugging)")
    ap.add_argument("--poses", type=int, default=None, help="override poses per GPU (debugging)")
    ap.add_argument("--extra", default=None, help="further configs measured in the same run, comma separated (default: c3, plus c4 at 8 GPUs; 'none')")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-per-frame", action="store_true", help="skip the per-frame call-pattern leg")
    ap.add_argument("--gather", default="p2p", choices=["p2p", "nccl"], help="N>1: how the clouds are exchanged")
    ap.add_argument("--gather-chunks", type=int, default=None, help="N>1: pose chunks per rank for the overlapped all-gather (default: 2 / 4 / 6 at 2 / 4 / 8 GPUs)")
    ap.add_argument("--gather-copy-self", action="store_true", help="N>1: compact into private buffers and let the exchange kernel copy to the own gather buffer too (round-1 behaviour)")
    ap.add_argument("--gather-ramp", type=int, default=None, help="N>1: first chunk = regular chunk / ramp")
    ap.add_argument("--gather-taper", type=int, default=3, help="N>1: last chunk = regular chunk / taper (its exchange is not hidden behind a traversal)")
    ap.add_argument("--wire", type=int, default=0, help="N>1: compact wire format of the exchange (1 on; default off)")
    ap.add_argument("--push-tile", type=int, default=None, help="N>1: bytes per stage of the TMA exchange kernel (2048 ... 16384)")
    ap.add_argument("--push-mode", type=int, default=None, help="N>1: exchange kernel, 0 = vector loads / stores, 1 = TMA bulk copies")
    ap.add_argument("--push-blocks", type=int, default=None, help="N>1: blocks per target of the exchange kernel")
    ap.add_argument("--e2e-chunk", type=int, default=None, help="poses per chunk of the pipelined e2e path")
    ap.add_argument("--e2e-inflight", type=int, default=2, help="host threads (each with its own engine context) issuing e2e trajectory calls")
    ap.add_argument("--no-numa-bind", action="store_true", help="N>1: do not pin each rank to its GPU's NUMA node")
    ap.add_argument("--node-format", type=int, default=None, help="0 = 64 B float node records, 1 = 32 B 16-bit records, 2 = 64 B paired records")
    ap.add_argument("--l2-persist", type=int, default=None, help="percent of the max persisting-L2 set-aside reserved for the BVH window (0 = off)")
    ap.add_argument("--variant", type=int, default=None, help="traversal kernel variant (lrc_set_option)")
    ap.add_argument("--tune", type=int, default=None, help="format-2 kernel tuning bits (lrc_set_option)")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--ref-frames", type=int, default=0, help="--impl reference: frames per step (0 = as many as fit --ref-budget)")
    ap.add_argument("--ref-budget", type=float, default=100.0, help="--impl reference: seconds of CPU work for the timed steps")
    args = ap.parse_args()
    if args.gather_chunks is None:
        n = max(args.gpus, int(os.environ.get("WORLD_SIZE", "1")))
        args.gather_chunks = 2 if n <= 2 else 4 if n <= 4 else 6          # measured: profiles/r02h_scaling.md
    if args.push_blocks is None and max(args.gpus, int(os.environ.get("WORLD_SIZE", "1"))) >= 8:
        args.push_blocks = 96
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus != world and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
