#!/usr/bin/env python
"""Throughput of the rows around the hot path on one B200, against the HBM copy peak and -- on a bounded
sample -- against the reference's own CPU implementation of each step (numpy / scikit-learn / the oracle's planner port).

    python tools/aux_bench.py [--out profiles/r01_aux_rows]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import lrc_b200 as lrc  # noqa: E402


def ev_time(fn, reps=10, flush=None):
    ts = []
    for k in range(reps + 2):
        if flush is not None:
            flush.fill_(k & 255)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if k >= 2:
            ts.append(a.elapsed_time(b))
    return float(np.mean(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_aux_rows"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak, _ = bench.measured_peak_gbs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    w, mesh, poses, intr = bench.make_workload(lrc, "c2", 1)
    eng = lrc.RaycastEngineGPU(device=0)
    ctx = eng.ctx
    scan = eng.simulate(poses, intr, mesh)
    M = scan.num_points
    rows = []

    # ---- f-4 scan statistics ----
    ms = ev_time(lambda: lrc.frame_statistics(ctx, scan), flush=flush)          # includes the tiny D2H of P records
    from lrc_b200 import _native as nat
    out = torch.empty((scan.num_frames, 40), dtype=torch.uint8, device=dev)
    ms_k = ev_time(lambda: nat.check(ctx._h, ctx._lib.lrc_frame_statistics(ctx._h, C.c_void_p(scan.points.data_ptr()), C.c_void_p(scan.incident.data_ptr()),
                                                                             C.c_void_p(scan.frame_offset.data_ptr()), scan.num_frames, C.c_void_p(out.data_ptr()), None)), flush=flush)
    host = scan.numpy()
    t0 = time.perf_counter()
    nfr = 5
    from oracle import post_oracle as po
    for p in range(nfr):
        a, b = host["frame_offset"][p], host["frame_offset"][p + 1]
        po.scan_quality(host["points"][a:b], host["incident"][a:b], 128000, 900.0)
    cpu_ms = (time.perf_counter() - t0) / nfr * scan.num_frames * 1e3
    rows.append({"row": "f-4 ScanQuality sums (lrc_frame_statistics)", "units": f"{M / 1e6:.1f} M points, {scan.num_frames} frames",
                 "gpu_ms": round(ms_k, 4), "alg_bytes": 20 * M, "gbs": round(20 * M / ms_k / 1e6, 1), "frac_hbm": round(20 * M / ms_k / 1e6 / peak, 3),
                 "cpu_ms_extrapolated": round(cpu_ms, 1), "cpu_what": "numpy expressions of s3dis_simulator.py:276-284, 1 core, 5 frames timed"})

    # ---- f-2 PLY records ----
    rec = torch.empty(19 * M, dtype=torch.uint8, device=dev)
    ms_k = ev_time(lambda: nat.check(ctx._h, ctx._lib.lrc_pack_ply_records(ctx._h, C.c_void_p(scan.points.data_ptr()), C.c_void_p(scan.label.data_ptr()), None, None,
                                                                             0x7F7F7F, M, C.c_void_p(rec.data_ptr()), None)), flush=flush)
    n_cpu = 200_000
    t0 = time.perf_counter()
    po.labeled_ply_bytes(host["points"][:n_cpu], np.full((n_cpu, 3), 127, np.uint8), (host["label"][:n_cpu] & 0xFFFF).astype(np.uint16),
                         (host["label"][:n_cpu] >> 16).astype(np.uint16))
    cpu_ms = (time.perf_counter() - t0) / n_cpu * M * 1e3
    rows.append({"row": "f-2 labelled-PLY records (lrc_pack_ply_records)", "units": f"{M / 1e6:.1f} M points",
                 "gpu_ms": round(ms_k, 4), "alg_bytes": 35 * M, "gbs": round(35 * M / ms_k / 1e6, 1), "frac_hbm": round(35 * M / ms_k / 1e6 / peak, 3),
                 "cpu_ms_extrapolated": round(cpu_ms, 1), "cpu_what": "per-point struct.pack loop of s3dis_sim_scene.py:634-641, 200k points timed"})

    # ---- f-3 1-NN transfer ----
    rng = np.random.default_rng(0)
    n_ref = 1_000_000
    ref = host["points"][rng.choice(M, n_ref, replace=False)].astype(np.float64) + rng.normal(0, 0.005, (n_ref, 3))
    t0 = time.perf_counter()
    lt = lrc.LabelTransfer(ctx, ref, semantic=rng.integers(0, 13, n_ref), colors=rng.random((n_ref, 3)))
    torch.cuda.synchronize()
    build_ms = (time.perf_counter() - t0) * 1e3
    ms_k = ev_time(lambda: lt.query(scan.points), reps=5, flush=flush)
    from sklearn.neighbors import NearestNeighbors
    t0 = time.perf_counter()
    nb = NearestNeighbors(n_neighbors=1, algorithm="ball_tree").fit(ref)
    fit_s = time.perf_counter() - t0
    nq = 200_000
    t0 = time.perf_counter()
    nb.kneighbors(host["points"][:nq])
    cpu_ms = (time.perf_counter() - t0) / nq * M * 1e3
    rows.append({"row": "f-3 1-NN label/colour transfer (lrc_nn_query)", "units": f"{M / 1e6:.1f} M queries vs {n_ref / 1e6:.0f} M annotated points",
                 "gpu_ms": round(ms_k, 3), "index_build_ms_incl_upload": round(build_ms, 1), "Mqueries_s": round(M / ms_k / 1e3, 1),
                 "cpu_ms_extrapolated": round(cpu_ms, 1), "cpu_fit_s": round(fit_s, 2),
                 "cpu_what": "scikit-learn ball_tree kneighbors, 1 core, 200k queries timed (s3dis_sim_scene.py:413-417)"})

    # ---- f-1 planner ----
    for name, m in (("10 x 8 m room, 50k tris", lrc.synthetic.box_room(50_000, seed=0)), ("60 x 40 m floor, 5M tris", lrc.synthetic.floor_plan(5_000_000, seed=0))):
        b = lrc.room_bounds_of(m)
        np.random.seed(1)
        gen = lrc.AutoTrajectoryGenerator(device=0)
        gen.generate_optimal_trajectory(m, b, 20)                        # warm-up (allocations)
        np.random.seed(1)
        gen = lrc.AutoTrajectoryGenerator(device=0)
        t0 = time.perf_counter()
        ra = gen._analyze_room_layout(m, b)
        torch.cuda.synchronize()
        t_layout = time.perf_counter() - t0
        gen.room_analysis = ra
        t0 = time.perf_counter()
        cands = gen._generate_trajectory_candidates(40)
        t_cand = time.perf_counter() - t0
        rows.append({"row": f"f-1 planner ({name})", "units": f"{len(m.vertices) / 1e6:.2f} M vertices, {len(ra.free_space_points)} free samples, "
                     f"{len(ra.connectivity_graph[1])} edges, {len(cands)} candidates",
                     "layout_ms": round(t_layout * 1e3, 1), "candidates_ms": round(t_cand * 1e3, 1),
                     "cpu_what": "reference planner on the 6k-triangle fixture room (1459 free samples): 7.5 s on this container's CPU; its loops are O(cells x V) + O(n^2)"})
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out + ".jsonl", "w") as fh:
        for r in rows:
            fh.write(json.dumps(r) + "\n")
            print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
