"""CPU: host-side logic of bench.py that decides what the JSON line may claim -- the committed ncu captures must belong to the
kernel build the library runs by default, a capture of another size is scaled by the ray count and never silently reused for
another mesh or build, both arms emit the same `config` object."""
import argparse
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "indoor-point-cloud-datasets-controllable-generation-method-for-mobile-robots-3d-scene-perception_b200"


def _default_tag():
    """The build tag bench.build_tag() reports for an untouched context, read from the library's option defaults."""
    src = open(os.path.join(ROOT, PKG, "csrc", "common.cuh")).read()

    def default(name):
        m = re.search(r"int64_t\s+%s\s*=\s*(-?\d+)\s*;" % name, src)
        assert m, name
        return int(m.group(1))
    return "fmt%d-q%d-leaf%d-var%d-tune%d-wp%d-rpt%d-pers%d" % tuple(default(n) for n in (
        "opt_node_format", "opt_build_quality", "opt_leaf_size", "opt_variant", "opt_tune", "opt_warp_packet",
        "opt_rays_per_thread", "opt_persistent"))


def test_committed_captures_belong_to_the_default_kernel_build():
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    tag = _default_tag()
    assert {"c2", "c3", "c4", "sweep_1e7"} <= set(tj)
    for key, e in tj.items():
        assert e["build_tag"] == tag, (key, e["build_tag"], tag)
        for k in ("tris", "rays_per_launch", "inst_executed", "issue_active_pct", "l1tex_pct", "dram_pct", "simt_lanes",
                  "dram_bytes_read", "dram_bytes_write", "dram_bytes_per_launch", "lts_t_bytes", "l1tex_t_bytes", "source"):
            assert k in e, (key, k)
        assert e["dram_bytes_per_launch"] == e["dram_bytes_read"] + e["dram_bytes_write"]
        assert 0 < e["issue_active_pct"] < 100 and 0 < e["simt_lanes"] <= 32
        assert os.path.exists(os.path.join(ROOT, re.search(r"profiles/\S+\.md", e["source"]).group(0))), e["source"]


def test_capture_lookup_scales_by_rays_and_refuses_other_builds():
    import bench
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["c2"]
    tag = tj["build_tag"]
    cap, why = bench.load_capture("c2", tj["tris"], tj["rays_per_launch"], tag)
    assert why is None and cap["inst_executed"] == tj["inst_executed"]
    half, why = bench.load_capture("c2", tj["tris"], tj["rays_per_launch"] // 4, tag)       # pose chunks at N > 1
    assert why is None and abs(half["inst_executed"] * 4 - tj["inst_executed"]) <= 4 and "scaled" in half["source"]
    assert abs(half["dram_bytes_per_launch"] * 4 - tj["dram_bytes_per_launch"]) <= 4 and "kernel_ms_ncu" not in half
    assert bench.load_capture("c2", tj["tris"] + 1, tj["rays_per_launch"], tag)[0] is None
    assert bench.load_capture("c2", tj["tris"], tj["rays_per_launch"], tag.replace("tune2", "tune0"))[0] is None
    assert bench.load_capture("c1", 1, 1, tag)[0] is None


def test_both_arms_emit_the_same_config_object():
    import bench
    args = argparse.Namespace(workload="c2")
    w = dict(bench.WORKLOADS["c2"])
    for world in (1, 2, 8):
        a = bench.workload_config(args, w, 999912, 128000, world)
        b = bench.workload_config(args, dict(w), 999912, 128000, world)
        assert a == b and a["poses_total"] == 100 * world and "l2" in a and "workload" in a
    assert bench.host_threads() >= 1
