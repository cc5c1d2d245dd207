"""
CPU ORACLE for the steps after the ray-cast call -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.py for who may import it).

Restates, expression by expression, in numpy / pure Python:
  * the ScanQuality arithmetic of the frame loop           /root/reference/s3dis_simulator.py:276-284
  * S3DISSimScene.compute_statistics and its getters        /root/reference/containers/s3dis_sim_scene.py:157-179,228-247
  * S3DISSimScene._save_labeled_ply (per-point struct.pack) /root/reference/containers/s3dis_sim_scene.py:614-641
  * the labelled-PLY reader of the downstream consumer      /root/reference/lidar_net_bbox_visualizer.py:72-126

Pinned by tests/golden/post_*.{npz,ply}, produced by the reference's own classes/functions
(tests/golden/make_golden_post.py).
"""
from __future__ import annotations

import struct

import numpy as np


def scan_quality(points: np.ndarray, incident_angles: np.ndarray, total_points_per_scan: int, room_volume: float) -> dict:
    """s3dis_simulator.py:276-284, verbatim expressions (points float32 (M,3), incident float64 (M,))."""
    return dict(
        coverage_ratio=len(points) / total_points_per_scan,
        num_points=len(points),
        incident_angle_mean=np.mean(incident_angles) if len(incident_angles) > 0 else 0,
        incident_angle_std=np.std(incident_angles) if len(incident_angles) > 0 else 0,
        scan_density=len(points) / room_volume,
        range_mean=np.mean(np.linalg.norm(points, axis=1)) if len(points) > 0 else 0,
        range_std=np.std(np.linalg.norm(points, axis=1)) if len(points) > 0 else 0,
    )


def simulation_stats(qualities, simulation_time: float) -> dict:
    """containers/s3dis_sim_scene.py:228-247 with the getters of :157-179."""
    if not qualities:
        return dict(total_frames=0, total_points=0, average_coverage=0.0, average_scan_density=0.0,
                    average_incident_angle=0.0, average_range=0.0, simulation_time=0.0, frames_per_second=0.0)
    return dict(
        total_frames=len(qualities),
        total_points=sum(q["num_points"] for q in qualities),
        average_coverage=np.mean([q["coverage_ratio"] for q in qualities]),
        average_scan_density=np.mean([q["scan_density"] for q in qualities]),
        average_incident_angle=np.mean([q["incident_angle_mean"] for q in qualities]),
        average_range=np.mean([q["range_mean"] for q in qualities]),
        simulation_time=simulation_time,
        frames_per_second=len(qualities) / simulation_time if simulation_time > 0 else 0.0,
    )


def labeled_ply_bytes(points, colors, semantic_labels, instance_labels) -> bytes:
    """containers/s3dis_sim_scene.py:614-641: header lines, then per point '<fff', '<BBB', '<HH'."""
    out = [b"ply\n", b"format binary_little_endian 1.0\n", b"element vertex %d\n" % len(points),
           b"property float x\n", b"property float y\n", b"property float z\n",
           b"property uchar red\n", b"property uchar green\n", b"property uchar blue\n",
           b"property ushort sem\n", b"property ushort ins\n", b"end_header\n"]
    for i in range(len(points)):
        out.append(struct.pack("<fff", points[i, 0], points[i, 1], points[i, 2]))
        out.append(struct.pack("<BBB", colors[i, 0], colors[i, 1], colors[i, 2]))
        out.append(struct.pack("<HH", semantic_labels[i], instance_labels[i]))
    return b"".join(out)


def read_labeled_ply_labels(path):
    """lidar_net_bbox_visualizer.py:72-126: header until end_header, then per vertex skip 15 bytes, unpack 'HH'."""
    with open(path, "rb") as f:
        header = []
        while True:
            line = f.readline().decode("utf-8").strip()
            header.append(line)
            if line == "end_header":
                break
        n = 0
        for line in header:
            if line.startswith("element vertex"):
                n = int(line.split()[-1])
        sem, ins = [], []
        for _ in range(n):
            f.read(12 + 3)
            s, i = struct.unpack("HH", f.read(4))
            sem.append(s)
            ins.append(i)
    return np.array(sem, dtype=np.uint16), np.array(ins, dtype=np.uint16)
