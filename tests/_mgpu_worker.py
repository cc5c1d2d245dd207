"""torchrun worker for tests/test_gpu_multi.py: pose-sharded simulation + NCCL all-gather must reproduce the
single-GPU result bit for bit on every rank (noise ON: Philox is keyed on the global pose index)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import lrc_b200 as lrc
    from lrc_b200.distributed import simulate_sharded
    mesh = lrc.synthetic.box_room(target_tris=20000, seed=1)
    poses = lrc.poses_from_waypoints([lrc.Waypoint(1.5 + 0.6 * k, 3.2 + 0.1 * k, 1.0, 0.05 * k) for k in range(11)])
    engine = lrc.RaycastEngineGPU(device=local)
    for intr in (lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis(), lrc.Indoor8LineLidarIntrinsics(max_range=5.0)):
        noise = lrc.NoiseConfig.from_intrinsics(intr, seed=77, pose_index_base=1000)
        got = simulate_sharded(engine, poses, intr, mesh, noise=noise).numpy()
        ref = engine.simulate(poses, intr, mesh, noise=noise).numpy()
        for k in ref:
            assert np.array_equal(got[k], ref[k]), (dist.get_rank(), type(intr).__name__, k)
        assert got["frame_offset"][-1] > 10000
    # steady-state overlapped path (what bench.py times for N > 1): xyz|label blocks gathered chunk by chunk
    from lrc_b200.distributed import OverlappedShardedScan
    world, rank = dist.get_world_size(), dist.get_rank()
    poses12 = lrc.poses_from_waypoints([lrc.Waypoint(1.5 + 0.5 * k, 3.2 + 0.1 * k, 1.0, 0.05 * k) for k in range(6 * world)])
    intr = lrc.DualAxisLidarIntrinsics.create_blk2go_dual_axis()
    noise = lrc.NoiseConfig.from_intrinsics(intr, seed=5, pose_index_base=0)
    sl = lrc.shard_range(len(poses12), rank, world)
    local_noise = lrc.NoiseConfig(noise.angle_noise_std, noise.dropout_probability, 0.0, noise.seed, sl.start)
    run = OverlappedShardedScan(engine.ctx, poses12[sl.start:sl.stop], intr, local_noise, chunks=3)
    for _ in range(2):
        run.step()
    got = run.assemble_numpy()
    ref = engine.simulate(poses12, intr, mesh, noise=noise).numpy()
    assert np.array_equal(got["frame_offset"], ref["frame_offset"])
    assert np.array_equal(got["points"], ref["points"]) and np.array_equal(got["label"], ref["label"])
    # fused compaction + all-gather over NVLink peer memory (lrc_set_gather): no collective at all
    from lrc_b200.distributed import PeerGather
    pg = PeerGather(engine.ctx, cap_per_rank=6 * 64000, frames_per_rank=6)
    pg.enable()
    engine.ctx.set_option("gather_chunks", 3)
    # uniform chunks; a short first chunk and few exchange blocks; a short last chunk; both short
    # ... and the same through the TMA exchange kernel (push_mode 1)
    for ramp, taper, blocks, mode in ((1, 1, 16, 0), (2, 1, 3, 0), (1, 3, 16, 0), (4, 4, 8, 0), (1, 1, 16, 1), (2, 3, 3, 1), (1, 1, 1, 1), (1, 1, 7, 2)):
        if mode == 2:                                  # the TMA kernel with its smallest tile (2 KB per stage)
            engine.ctx.set_option("push_tile", 2048)
            mode = 1
        engine.ctx.set_option("push_mode", mode)
        engine.ctx.set_option("gather_ramp", ramp)
        engine.ctx.set_option("gather_taper", taper)
        engine.ctx.set_option("push_blocks", blocks)
        pg.buffer.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        for _ in range(2):
            local = engine.simulate(poses12[sl.start:sl.stop], intr, noise=local_noise).numpy()
        pg.synchronize()
        got = pg.assemble_numpy(poses_all=poses12)
        assert np.array_equal(got["frame_offset"], ref["frame_offset"]), ramp
        assert np.array_equal(got["points"], ref["points"]) and np.array_equal(got["label"], ref["label"]), ramp
        # the reference's frame is (points, incident angles): the angles are recomputed on arrival, bit for bit
        assert np.array_equal(got["incident"], ref["incident"]), ramp
        dist.barrier()
    # compaction straight into this rank's region of its own gather buffer (no copy to itself), both exchange kernels
    for mode in (0, 1):
        engine.ctx.set_option("push_mode", mode)
        pg.buffer.zero_()
        torch.cuda.synchronize()
        dist.barrier()
        bufs = pg.local_out(sl.stop - sl.start)
        engine.ctx.scan(poses12[sl.start:sl.stop], intr, local_noise, bufs=bufs)
        pg.synchronize()
        got = pg.assemble_numpy(poses_all=poses12)
        for k in ("frame_offset", "points", "label", "incident"):
            assert np.array_equal(got[k], ref[k]), ("local_out", mode, k)
        dist.barrier()
    engine.ctx.set_option("push_mode", 0)
    # a scan of more frames than this rank's offset region holds must be refused, not spill into the neighbour's region
    try:
        engine.simulate(poses12[:7], lrc.Indoor8LineLidarIntrinsics(max_range=5.0))      # 7 x 16000 rays fit, 7 + 1 offsets do not
        raise AssertionError("expected LRC_ERR_CAPACITY")
    except RuntimeError as e:
        assert "frame_capacity" in str(e), e
    pg.disable()
    engine.ctx.set_option("push_tile", 16384)
    engine.ctx.set_option("gather_ramp", 1)
    engine.ctx.set_option("gather_taper", 1)
    a, b = ref["frame_offset"][sl.start], ref["frame_offset"][sl.stop]
    assert np.array_equal(local["points"], ref["points"][a:b]) and np.array_equal(local["incident"], ref["incident"][a:b])
    pg.close()
    dist.barrier()
    # compact wire format: t | label | ray index travel, every rank rebuilds the other ranks' points on arrival
    # (noisy dual-axis sensor and a single-axis sensor whose frame is not a multiple of the block size)
    nfs = [len(lrc.shard_range(len(poses12), r, world)) for r in range(world)]
    for sensor, nz_all in ((intr, noise), (lrc.Indoor8LineLidarIntrinsics(max_range=5.0, horizontal_res=999), None)):
        refw = engine.simulate(poses12, sensor, mesh, noise=nz_all).numpy()
        n_ray = lrc.rays_per_frame(sensor)
        pgw = PeerGather(engine.ctx, cap_per_rank=max(nfs) * n_ray, frames_per_rank=max(nfs))
        pgw.enable(wire=True, poses_all=poses12, frames_per_rank=nfs)
        ln = None if nz_all is None else lrc.NoiseConfig(nz_all.angle_noise_std, nz_all.dropout_probability, 0.0, nz_all.seed, sl.start)
        for chunks in (3, 1, 2):
            engine.ctx.set_option("gather_chunks", chunks)
            pgw.buffer[: pgw.o_lab].zero_()                       # wipe every rank's xyz: what is compared below was rebuilt in this pass
            torch.cuda.synchronize()
            dist.barrier()
            engine.ctx.scan(poses12[sl.start:sl.stop], sensor, ln, bufs=pgw.local_out(sl.stop - sl.start))
            pgw.synchronize()
            got = pgw.assemble_numpy(frames_per_rank=nfs, poses_all=poses12)
            for k in ("frame_offset", "points", "label", "incident"):
                assert np.array_equal(got[k], refw[k]), ("wire", type(sensor).__name__, chunks, k)
            dist.barrier()
        pgw.close()
        dist.barrier()
    engine.ctx.set_option("gather_chunks", 4)
    if dist.get_rank() == 0:
        print("MGPU_OK world=%d" % dist.get_world_size())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
