"""Host-side logic of the N > 1 path on CPU: contiguous pose sharding + the variable-length all-gather, with
world_size-2/3 gloo process groups.  Each rank's local records come from the oracle; the gathered result must
equal the single-process result bit for bit (the property the 8-GPU run must have)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_frames(pose_indices, noise_base=0):
    """Compacted records of the given frames of a fixed small trajectory, from the CPU oracle."""
    import lrc_b200 as lrc
    from oracle import oracle as orc
    mesh = lrc.synthetic.box_room(target_tris=3000, seed=5)
    scene = orc.OracleScene((mesh.vertices, mesh.triangles))
    intr = lrc.DualAxisLidarIntrinsics(point_rate=6000, scan_duration=0.1, num_vertical_lines=6, max_range=4.0)
    wps = [lrc.Waypoint(2.0 + 0.9 * k, 3.0, 1.0, 0.2 * k) for k in range(7)]
    recs = {k: [] for k in ("points", "incident", "prim_id", "label", "ray_idx")}
    counts = []
    for p in pose_indices:
        pose = wps[p].to_pose_matrix()
        rays, keep = orc.gen_rays_dual_axis(pose, orc.dual_params(intr), seed=9, pose_idx=noise_base + p, compact=False)
        t, pid = scene.cast_rays(rays)
        fr = orc.epilogue_c(rays, t, pid, center=pose[:3, 3], max_range=intr.max_range, tri_label=mesh.triangle_labels,
                            keep=keep.astype(np.uint8))
        recs["points"].append(fr.points); recs["incident"].append(fr.incident)
        recs["prim_id"].append(fr.prim_id.view(np.int32)); recs["label"].append(fr.label.view(np.int32))
        recs["ray_idx"].append(fr.ray_idx.view(np.int32))
        counts.append(len(fr.points))
    dt = {"points": np.float32, "incident": np.float64, "prim_id": np.int32, "label": np.int32, "ray_idx": np.int32}
    out = {}
    for k, v in recs.items():
        arr = np.concatenate(v) if v else np.zeros((0, 3) if k == "points" else (0,), dt[k])
        out[k] = torch.from_numpy(np.ascontiguousarray(arr))
    return out, torch.tensor(counts, dtype=torch.int64)


def _worker(rank, world, port, total, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import lrc_b200 as lrc
        from lrc_b200.distributed import allgather_clouds
        sl = lrc.shard_range(total, rank, world)
        local, counts = _oracle_frames(list(sl))
        got = allgather_clouds(local, counts, total).numpy()
        q.put((rank, {k: v.copy() for k, v in got.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 7), (3, 7), (2, 1)])
def test_sharded_gather_equals_single_process(world, total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref, counts = _oracle_frames(list(range(total)))
    ref_off = np.concatenate([[0], np.cumsum(counts.numpy())])
    assert ref_off[-1] > 100
    for r in range(world):
        got = results[r]
        assert np.array_equal(got["frame_offset"], ref_off)
        assert np.array_equal(got["points"], ref["points"].numpy())
        assert np.array_equal(got["incident"], ref["incident"].numpy())
        for k in ("prim_id", "label", "ray_idx"):
            assert np.array_equal(got[k], ref[k].numpy().view(np.uint32)), k


def test_shard_ranges_partition_the_trajectory():
    import lrc_b200 as lrc
    for total in (0, 1, 7, 100, 500):
        for world in (1, 2, 3, 4, 8):
            parts = [lrc.shard_range(total, r, world) for r in range(world)]
            assert sum(len(p) for p in parts) == total
            flat = [i for p in parts for i in p]
            assert flat == list(range(total))                     # contiguous, ordered, disjoint
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
