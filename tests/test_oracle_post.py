"""CPU: the post-cast oracle (oracle/post_oracle.py) against fixtures produced by the reference's own classes
(tests/golden/make_golden_post.py), plus the host-side halves of lrc_b200.post that need no GPU."""
import os

import numpy as np

from conftest import GOLDEN

Q_KEYS = ("coverage_ratio", "num_points", "incident_angle_mean", "incident_angle_std", "scan_density", "range_mean", "range_std")
S_KEYS = ("total_frames", "total_points", "average_coverage", "average_scan_density", "average_incident_angle",
          "average_range", "simulation_time", "frames_per_second")


def _post_oracle():
    from oracle import post_oracle
    return post_oracle


def test_scan_quality_matches_reference_dataclass(golden):
    po = _post_oracle()
    g = golden("post_stats.npz")
    tot, vol = int(g["total_points_per_scan"]), float(g["room_volume"])
    qs = []
    for i in range(3):
        q = po.scan_quality(g[f"frame{i}/points"], g[f"frame{i}/incident"], tot, vol)
        qs.append(q)
        want = g[f"frame{i}/quality"]
        got = np.array([q[k] for k in Q_KEYS], dtype=np.float64)
        assert np.array_equal(got, want), (i, got, want)          # same numpy expressions -> same bits
    s = po.simulation_stats(qs, 2.5)
    assert np.array_equal(np.array([s[k] for k in S_KEYS], dtype=np.float64), g["stats"])
    assert qs[1]["num_points"] == 0 and qs[1]["range_mean"] == 0           # the empty frame


def test_labeled_ply_bytes_match_reference_writer(golden):
    po = _post_oracle()
    g = golden("post_stats.npz")
    want = open(os.path.join(GOLDEN, "post_labeled.ply"), "rb").read()
    got = po.labeled_ply_bytes(g["ply/points"], g["ply/colors"], g["ply/sem"], g["ply/ins"])
    assert got == want
    sem, ins = po.read_labeled_ply_labels(os.path.join(GOLDEN, "post_labeled.ply"))
    assert np.array_equal(sem, g["ply/sem"]) and np.array_equal(ins, g["ply/ins"])


def test_host_header_and_reader(lrc, golden, tmp_path):
    g = golden("post_stats.npz")
    ref = open(os.path.join(GOLDEN, "post_labeled.ply"), "rb").read()
    n = len(g["ply/points"])
    hdr = lrc.post.ply_header(n)
    assert ref.startswith(hdr) and len(ref) == len(hdr) + 19 * n
    rec = lrc.read_labeled_ply(os.path.join(GOLDEN, "post_labeled.ply"))
    assert np.array_equal(rec["points"].view(np.uint32), g["ply/points"].view(np.uint32))
    assert np.array_equal(rec["colors"], g["ply/colors"])
    assert np.array_equal(rec["semantic_labels"], g["ply/sem"]) and np.array_equal(rec["instance_labels"], g["ply/ins"])
    rgb = lrc.post.pack_rgb(g["ply/colors"])
    assert rgb.dtype == np.uint32 and int(rgb[5]) == int(g["ply/colors"][5, 0]) | int(g["ply/colors"][5, 1]) << 8 | int(g["ply/colors"][5, 2]) << 16


def test_simulation_stats_host_fixes_fps(lrc):
    qs = [lrc.ScanQuality(0.5, 8000, 40.0, 10.0, 33.3, 5.0, 1.0), lrc.ScanQuality(0.25, 4000, 50.0, 12.0, 16.7, 7.0, 2.0)]
    s = lrc.simulation_stats(qs, 0.5)
    assert s.total_frames == 2 and s.total_points == 12000 and s.frames_per_second == 4.0
    assert s.average_incident_angle == 45.0 and s.average_range == 6.0
    assert lrc.simulation_stats([], 1.0).frames_per_second == 0.0
