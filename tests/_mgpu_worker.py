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
    dist.barrier()
    if dist.get_rank() == 0:
        print("MGPU_OK world=%d" % dist.get_world_size())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
