/*
 * lrc.h -- C ABI of the B200-native LiDAR ray-casting engine (liblrc.so, sm_100a).
 *
 * This is the drop-in boundary for the ONE hot path of the reference:
 *     RaycastEngine*.lidar_intersect_mesh / rays_intersect_mesh, called once per waypoint from
 *     S3DISSimulator.run_simulation            (reference s3dis_simulator.py:254-288, :261)
 * The reference has no FFI of its own (pure Python over Open3D/Embree); every entry point below
 * names the reference interface it replaces.  All citations are relative to the reference tree.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no C++/torch types cross the boundary.
 *   - Every function returns 0 on success and a negative lrc_status on failure; the message is
 *     available from lrc_last_error(ctx) (or lrc_last_error(NULL) for lrc_create failures).
 *   - One lrc_ctx per (process, GPU).  A context is NOT thread-safe; callers serialise per context.
 *   - Pointers are DEVICE pointers on ctx's GPU unless the parameter name starts with `h_`.
 *   - All work is enqueued on the caller's `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream).  Outputs are valid once that stream has been synchronised.  The caller
 *     owns every buffer passed in; the context owns the BVH and its scratch space.
 *   - Ray layout: N x 6 float32 row-major [ox oy oz dx dy dz]; directions are used AS GIVEN
 *     (not normalised) by the intersector, t is in units of |d|  (raycast_engine_cpu.py:50-51).
 *   - Miss: t_hit = +inf, prim_id = LRC_MISS_ID.
 *   - Closest hit: smallest t; equal t -> smallest original triangle index.  Two-sided triangles,
 *     t >= 0.  The float32 Moller-Trumbore operation order is fixed (DESIGN.md "intersection spec").
 *   - Triangle label: uint32 = semantic | instance << 16 (dtype contract of
 *     containers/s3dis_sim_scene.py:581-582,630-631).
 */
#ifndef LRC_H
#define LRC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRC_MISS_ID 0xFFFFFFFFu
#define LRC_ABI_VERSION 1

typedef struct lrc_ctx lrc_ctx;

typedef enum {
    LRC_OK = 0,
    LRC_ERR_INVALID = -1,   /* bad argument (NULL pointer, negative size, index out of range ...) */
    LRC_ERR_CUDA = -2,      /* a CUDA runtime call failed; message holds cudaGetErrorString */
    LRC_ERR_NO_MESH = -3,   /* a cast/scan entry point was called before lrc_set_mesh */
    LRC_ERR_CAPACITY = -4,  /* an output buffer is too small / BVH deeper than the traversal stack */
    LRC_ERR_NO_DEVICE = -5  /* no usable CUDA device: the engine has no CPU fallback */
} lrc_status;

/* Single-axis sensor == Indoor8LineLidarIntrinsics + IndoorLidar.get_rays
 * (lidar/lidar_intrinsics.py:214-289, lidar/indoor_lidar.py:27-53,56-131). */
typedef struct {
    int32_t H;                    /* number of scan lines (len(vertical_degrees), or vertical_res) */
    int32_t W;                    /* horizontal_res */
    const double* h_vertical_deg; /* HOST, H entries, degrees; NULL -> uniform-fov table below   */
    double fov_up_deg;            /* used only when h_vertical_deg == NULL (indoor_lidar.py:56-91) */
    double fov_down_deg;
    double max_range;             /* metres; strict '<' on the f64 distance (raycast_engine_cpu.py:95-97) */
} lrc_single_axis;

/* Dual-axis sensor == DualAxisLidarIntrinsics + DualAxisLidar.get_multi_line_rays
 * (lidar/lidar_intrinsics.py:28-66,152-186, lidar/indoor_lidar.py:224-296). */
typedef struct {
    int32_t num_lines;            /* num_vertical_lines */
    int32_t points_per_line;      /* int(point_rate*scan_duration) // num_vertical_lines */
    double theta_min, theta_max;  /* theta_range, radians */
    double swing_amplitude;       /* radians */
    double swing_frequency;
    double max_range;
} lrc_dual_axis;

/* Noise model.  The reference draws from the global numpy stream (indoor_lidar.py:270-272,292-294);
 * the engine uses Philox4x32-10 keyed on (seed, pose index, ray index) so that results do not
 * depend on launch geometry or on the number of GPUs.  NULL or all-zero = noise disabled. */
typedef struct {
    double angle_noise_std;       /* rad, added to phi and theta after the clip (dual-axis only) */
    double dropout_probability;   /* ray kept iff u > p, BEFORE casting (dual-axis only) */
    double range_noise_std;       /* metres, added to t before the point is reconstructed (both) */
    uint64_t seed;
    uint64_t pose_index_base;     /* global index of poses[0]: rank r of a sharded run passes its offset */
} lrc_noise;

/* Compacted, ray-ordered output of a scan (structure of arrays, caller-allocated, device).
 * Frame p occupies [frame_offset[p], frame_offset[p+1]).  Any array pointer except xyz and
 * frame_offset may be NULL (that attribute is then not produced). */
typedef struct {
    float* xyz;              /* capacity x 3 float32                 == points  (raycast_engine_cpu.py:111) */
    double* incident_deg;    /* capacity      float64, degrees       == incident_angles (:100-107) */
    uint32_t* prim_id;       /* capacity      original triangle index */
    uint32_t* label;         /* capacity      tri_label[prim_id] (0 when the mesh has no labels) */
    uint32_t* ray_idx;       /* capacity      index of the ray inside its frame (before dropout) */
    int64_t* frame_offset;   /* P + 1 */
    int64_t capacity;        /* in points; P * rays_per_pose always suffices */
} lrc_out;

typedef struct {
    uint64_t rays;           /* rays traversed (after dropout) */
    uint64_t nodes_visited;  /* BVH node records fetched (64 B each) */
    uint64_t tris_tested;    /* triangle records fetched (48 B each) */
    uint64_t hits;           /* rays with a finite t */
} lrc_counters_t;

typedef struct {
    int64_t num_tris;
    int64_t num_nodes;
    int32_t max_depth;
    int32_t reserved;
    float scene_min[3];
    float scene_max[3];
    float box_pad;           /* absolute padding applied to every leaf box */
    float sah_cost;          /* surface-area-heuristic cost of the built tree (diagnostic) */
    int64_t bytes_nodes;
    int64_t bytes_tris;
} lrc_bvh_info;

/* ---- lifetime ------------------------------------------------------------------------------- */
int lrc_abi_version(void);
/* == RaycastEngineGPU.__init__ (raycast_engine/raycast_engine_gpu_simple.py:18-20). */
int lrc_create(int device, lrc_ctx** out);
void lrc_destroy(lrc_ctx* ctx);
const char* lrc_last_error(const lrc_ctx* ctx);

/* ---- scene ---------------------------------------------------------------------------------- */
/* == o3d.t.geometry.RaycastingScene() + add_triangles(from_legacy(mesh))
 *    (raycast_engine_cpu.py:46-47).  verts are float32 (the caller rounds the legacy float64
 *    vertices exactly as from_legacy does).  Builds the LBVH (Morton codes + radix sort + Karras
 *    hierarchy + refit) on the GPU and keeps it until the next lrc_set_mesh / lrc_destroy.
 *    tri_label may be NULL. */
int lrc_set_mesh(lrc_ctx* ctx, const float* verts, int64_t V, const int32_t* tris, int64_t T,
                 const uint32_t* tri_label, void* stream);
int lrc_bvh_get_info(lrc_ctx* ctx, lrc_bvh_info* h_info);

/* ---- rays_intersect_mesh -------------------------------------------------------------------- */
/* == RaycastingScene.cast_rays -> "t_hit", "primitive_ids" (raycast_engine_cpu.py:51-53). */
int lrc_cast_rays(lrc_ctx* ctx, const float* rays, int64_t N, float* t_hit, uint32_t* prim_id, void* stream);
/* Exhaustive ray x triangle reference kernel on the GPU (validation of the BVH path). */
int lrc_cast_rays_bruteforce(lrc_ctx* ctx, const float* rays, int64_t N, float* t_hit, uint32_t* prim_id, void* stream);
/* == RaycastEngineCPU.rays_intersect_mesh (raycast_engine_cpu.py:24-73): cast, d^ = d/|d| and
 *    p = o + d^*t in float32, ordered compaction by the hit mask.  out->frame_offset has 2 entries. */
int lrc_rays_intersect(lrc_ctx* ctx, const float* rays, int64_t N, lrc_out* out, void* stream);
/* == RaycastEngineCPU.lidar_intersect_mesh (raycast_engine_cpu.py:75-111) for a duck-typed lidar
 *    whose rays were produced elsewhere: cast + range filter around h_center + incident angle. */
int lrc_scan_rays(lrc_ctx* ctx, const float* rays, int64_t N, const double* h_center, double max_range,
                  lrc_out* out, void* stream);

/* ---- lidar_intersect_mesh, batched over a trajectory ---------------------------------------- */
/* P poses, each a row-major 4x4 float64 matrix (Waypoint.to_pose_matrix,
 * trajectory/trajectory_generator.py:30-44), DEVICE memory.  Rays are generated in-kernel and
 * never written to HBM.  Per frame this is RaycastEngineCPU.lidar_intersect_mesh. */
int lrc_scan_single_axis(lrc_ctx* ctx, const double* poses, int64_t P, const lrc_single_axis* h_sensor,
                         const lrc_noise* h_noise, lrc_out* out, void* stream);
int lrc_scan_dual_axis(lrc_ctx* ctx, const double* poses, int64_t P, const lrc_dual_axis* h_sensor,
                       const lrc_noise* h_noise, lrc_out* out, void* stream);

/* ---- the same, with HOST buffers (the reference-facing call: numpy in, numpy out) -------------- */
/* h_poses: P x 16 float64 on the host.  h_out: an lrc_out whose pointers are HOST memory (page-locked memory makes
 * the copies asynchronous; pageable memory works, slower).  The trajectory is cut into chunks of `chunk_poses` poses
 * (0 = automatic: a first chunk of ~0.5M rays, each following chunk twice as long, so that the PCIe copy -- the long
 * pole -- starts early and then moves few, large pieces); every chunk's kernels are enqueued at once on an internal
 * stream and each chunk's compacted records are copied back on a second stream as soon as that chunk is finished, so
 * the PCIe transfer overlaps the traversal of later chunks.  Synchronous: on return every output is on the host and *h_num_points is set.
 * These two calls do NOT use the caller's stream. */
int lrc_scan_single_axis_host(lrc_ctx* ctx, const double* h_poses, int64_t P, const lrc_single_axis* h_sensor,
                              const lrc_noise* h_noise, lrc_out* h_out, int64_t chunk_poses, int64_t* h_num_points);
int lrc_scan_dual_axis_host(lrc_ctx* ctx, const double* h_poses, int64_t P, const lrc_dual_axis* h_sensor,
                            const lrc_noise* h_noise, lrc_out* h_out, int64_t chunk_poses, int64_t* h_num_points);
/* == lrc_set_mesh with HOST arrays (uploads, then builds; synchronous). */
int lrc_set_mesh_host(lrc_ctx* ctx, const float* h_verts, int64_t V, const int32_t* h_tris, int64_t T,
                      const uint32_t* h_tri_label);

/* ---- multi-GPU: all-gather of the compacted clouds over NVLink peer memory, overlapped with traversal -------- */
/* The reference has no multi-process path; this is the exchange step of the pose-sharded run.
 * A peer buffer is plain device memory whose CUDA IPC handle other ranks open; lrc_set_gather names up to 16
 * target buffers (this GPU's own and the peers' mapped ones).  While targets are set, every scan cuts the trajectory
 * into pose chunks ("gather_chunks"); as soon as a chunk has been compacted locally, an exchange kernel copies that
 * chunk's xyz and label slices into EVERY target at point_base + position with 16-byte vector stores, and the chunk's
 * frame offsets at frame_base + frame (+ the closing entry at frame_base + P).  Slice bounds are read on the device
 * (no host round trip); the stores to peers travel over NVLink while the next chunk is being traversed.  point_base
 * should be a multiple of 4 (keeps source and destination congruent modulo 16 bytes).  Callers synchronise their
 * stream and then barrier across ranks before reading.  The scan's lrc_out must carry a label array. */
#define LRC_MAX_GATHER_TARGETS 16
typedef struct { unsigned char bytes[64]; } lrc_ipc_handle;
typedef struct {
    int32_t n_targets;
    int32_t frame_capacity;  /* frame-offset slots of this rank's region (a scan of P frames writes P + 1); 0 = not checked */
    float* xyz[LRC_MAX_GATHER_TARGETS];
    uint32_t* label[LRC_MAX_GATHER_TARGETS];
    int64_t* frame_offset[LRC_MAX_GATHER_TARGETS];
    int64_t point_base;      /* first point slot of this rank's region inside every target */
    int64_t frame_base;      /* first frame-offset slot of this rank's region */
    int64_t capacity;        /* point slots available to this rank's region */
} lrc_gather;
int lrc_peer_buffer_create(lrc_ctx* ctx, int64_t bytes, void** d_ptr, lrc_ipc_handle* h_handle);
int lrc_peer_buffer_open(lrc_ctx* ctx, const lrc_ipc_handle* h_handle, void** d_ptr);
int lrc_peer_buffer_close(lrc_ctx* ctx, void* d_ptr);
int lrc_peer_buffer_destroy(lrc_ctx* ctx, void* d_ptr);
int lrc_set_gather(lrc_ctx* ctx, const lrc_gather* h_targets /* NULL or n_targets == 0: off */);
/* Compact wire format of the exchange (optional; call after lrc_set_gather, which switches it off again).  A gathered
 * point is then t | label | ray index (12 B) instead of xyz | label (16 B): every rank knows every pose and the sensor, so
 * it REBUILDS the points of the other ranks' frames on arrival -- regenerate the ray (same float64 table arithmetic as the
 * scan, one rounding to float32), p = o + (d / |d|) * t with the scan epilogue's own float32 operations -- bit-identical to
 * what the producing rank computed, for 25 % less NVLink traffic (the all-gather is ingress-bound from 4 GPUs on).  The
 * producer marks its progress in every target (ready[target][self] = scan number << 32 | frames pushed so far, release at
 * system scope, after the chunk's bulk copies have completed); the receiver waits for that word (acquire, with a time-out
 * that fails the scan instead of hanging) and rebuilds chunk by chunk on its own stream, inside the scan call.
 * All ranks must issue the same sequence of scans while this is enabled (the scan number tags the progress words), and a
 * rank's noise pose_index_base must be (global base + rank_pose0[self]).  The scan's lrc_out must NOT carry incident_deg
 * (the scratch slot of the angle carries t; lrc_incident_angles recomputes the angles of any gathered cloud). */
typedef struct {
    int32_t enabled;                                 /* 0: off */
    int32_t self;                                    /* index of this rank among the gather targets */
    float* t[LRC_MAX_GATHER_TARGETS];                /* per target: hit distances, indexed like label[] */
    uint32_t* ray_idx[LRC_MAX_GATHER_TARGETS];       /* per target: ray index inside the frame, indexed like label[] */
    int64_t* ready[LRC_MAX_GATHER_TARGETS];          /* per target: n_targets progress words; word r is written by rank r */
    const double* all_poses;                         /* device: poses of ALL ranks, rank-major, (sum of rank_frames) x 16 float64 */
    int64_t rank_pose0[LRC_MAX_GATHER_TARGETS];      /* first pose of every rank inside all_poses */
    int64_t rank_frames[LRC_MAX_GATHER_TARGETS];     /* frames every rank scans per call */
    int64_t rank_point_base[LRC_MAX_GATHER_TARGETS]; /* first point slot of every rank's region */
    int64_t rank_frame_base[LRC_MAX_GATHER_TARGETS]; /* first frame-offset slot of every rank's region */
} lrc_gather_wire;
int lrc_set_gather_wire(lrc_ctx* ctx, const lrc_gather_wire* h_wire /* NULL or enabled == 0: off */);

/* ---- get_rays ------------------------------------------------------------------------------- */
/* == IndoorLidar.get_rays (indoor_lidar.py:27-53): rays H*W x 6 float32, index j*W + i. */
int lrc_gen_rays_single_axis(lrc_ctx* ctx, const double* poses, int64_t P, const lrc_single_axis* h_sensor,
                             float* rays, void* stream);
/* == DualAxisLidar.get_rays (indoor_lidar.py:311-319): dense table line*ppl + k plus keep flags
 *    (the caller drops rays whose flag is 0, as indoor_lidar.py:292-294 does). */
int lrc_gen_rays_dual_axis(lrc_ctx* ctx, const double* poses, int64_t P, const lrc_dual_axis* h_sensor,
                           const lrc_noise* h_noise, float* rays, uint8_t* keep, void* stream);

/* ---- after the cast: per-frame statistics and the labelled-PLY wire format ------------ */
/* == the ScanQuality sums of the frame loop (s3dis_simulator.py:276-284): per frame num_points, np.mean / np.std
 *    (population) of the float64 incident angles, and np.mean / np.std of np.linalg.norm(points, axis=1) -- float32
 *    norms taken from the WORLD ORIGIN, as the reference does.  Sums are carried in float64 in a fixed order
 *    (deterministic); coverage_ratio and scan_density are num_points divided by host constants.  incident_deg may be
 *    NULL (angle fields are then 0).  out: P records, device. */
typedef struct {
    int64_t num_points;
    double incident_mean, incident_std;
    double range_mean, range_std;
} lrc_frame_stats;
int lrc_frame_statistics(lrc_ctx* ctx, const float* xyz, const double* incident_deg, const int64_t* frame_offset,
                         int64_t P, lrc_frame_stats* out, void* stream);
/* == the incident-angle lines of RaycastEngineCPU.lidar_intersect_mesh (raycast_engine_cpu.py:100-107) for points that are
 *    already known: incident_deg[i] = degrees(arccos(|(p_i - c)_z / ||p_i - c|||)), float64, c = poses[frame(i)][:3, 3],
 *    frame(i) from frame_offset (P + 1 entries; frame f owns [frame_offset[f], frame_offset[f + 1])).  Bit-identical to the
 *    angles a scan produces for the same points; used after the multi-GPU exchange, which moves xyz + label only.
 *    xyz: M x 3 float32, poses: P x 16 float64, all device.  frame_offset may carry a base: xyz[0] is absolute point
 *    frame_offset[0]; M is an upper bound (a capacity), points at or beyond frame_offset[P] are left untouched. */
int lrc_incident_angles(lrc_ctx* ctx, const float* xyz, const int64_t* frame_offset, int64_t P, const double* poses, int64_t M,
                        double* incident_deg, void* stream);
/* == the vertex records S3DISSimScene._save_labeled_ply writes with struct.pack per point
 *    (containers/s3dis_sim_scene.py:634-641): M x 19 bytes, little endian, '<fff BBB HH' =
 *    x y z | red green blue | sem ins.  label = sem | ins << 16 (NULL: both 0, the reference's default labels,
 *    s3dis_sim_scene.py:584-594).  Colour: tri_rgb[prim_id[i]] (red | green << 8 | blue << 16) when both are given,
 *    else default_rgb (0x7F7F7F is the reference's default grey, (0.5 * 255).astype(uint8)).  out: 19 * M bytes,
 *    16-byte aligned, device.  The caller prepends the text header (:621-632) and writes the bytes. */
int lrc_pack_ply_records(lrc_ctx* ctx, const float* xyz, const uint32_t* label, const uint32_t* prim_id,
                         const uint32_t* tri_rgb, uint32_t default_rgb, int64_t M, uint8_t* out, void* stream);

/* ---- 1-nearest-neighbour label / colour transfer -------------------------------------------- */
/* == NearestNeighbors(n_neighbors=1, algorithm='ball_tree').fit(ref_pts) (scikit-learn, third party; reference
 *    containers/s3dis_sim_scene.py:413-417 and :536-538).  ref_pts: n x 3 float64, device (the annotated S3DIS points).
 *    Bins them on a uniform 3-D grid of `cell` metres (0 = automatic); kept until the next build. */
int lrc_nn_index_build(lrc_ctx* ctx, const double* ref_pts, int64_t n, double cell, void* stream);
/* == .kneighbors(points) + the gathers colours[indices], labels[indices] (:418-424).  query_xyz: M x 3 float32 (the
 *    frame's hit points; promoted to float64 like scikit-learn does).  out_index[i] = argmin_j of the float64 reduced
 *    distance ((dx*dx + dy*dy) + dz*dz), ties -> smaller j; -1 when the index is empty or the query is not finite.
 *    out_distance (may be NULL): Euclidean distance.  Up to two uint32 attribute tables of the annotated points
 *    (e.g. sem | ins << 16 and packed rgb) are gathered through the indices when both ref_attr_x and out_attr_x are
 *    given.  All arrays device. */
int lrc_nn_query(lrc_ctx* ctx, const float* query_xyz, int64_t M, int32_t* out_index, double* out_distance,
                 const uint32_t* ref_attr_a, uint32_t* out_attr_a, const uint32_t* ref_attr_b, uint32_t* out_attr_b,
                 void* stream);

/* ---- the caller that produces the poses: coverage-trajectory planner support --------------- */
/* == the vertex set AutoTrajectoryGenerator._is_point_inside_mesh scans for every query point
 *    (trajectory/auto_trajectory_generator.py:220-238), binned once on a 2-D grid of `cell` metres.
 *    verts: V x 3 float64 (the legacy mesh's vertices, NOT rounded to float32), device.  Kept until the next build. */
int lrc_collision_index_build(lrc_ctx* ctx, const double* verts, int64_t V, double cell, void* stream);
/* state[q] = 0  the robot cube [p - half, p + half] leaves h_bounds = {x_min, x_max, y_min, y_max, z_min, z_max}
 *               (_is_point_in_room_bounds, :204-217; h_bounds NULL = no bounds test)
 *          = 1  some vertex lies inside the closed cube (_is_point_inside_mesh, :220-238)
 *          = 2  free.                         pts: Q x 3 float64, device; state: Q bytes, device. */
int lrc_collision_query(lrc_ctx* ctx, const double* pts, int64_t Q, double half, const double* h_bounds,
                        uint8_t* state, void* stream);
/* == _build_connectivity_graph (:245-258) for grid samples (xs[ix], ys[iy], z) in the reference's x-major order:
 *    free_index[ix * ny + iy] = rank of the cell among the free ones (state == 2) or -1; CSR rows = free points,
 *    columns = the free points j != i with ||p_i - p_j|| <= max_dist, ascending.  `window` = how many grid steps
 *    around a cell are examined (ceil(max_dist / step) + 1 is always enough).  h_counts[0] = number of free points,
 *    h_counts[1] = number of edges.  row_ptr: capacity nx*ny + 1; col: capacity col_capacity (LRC_ERR_CAPACITY if
 *    smaller than the edge count, with h_counts already set).  All arrays device except h_counts.  Synchronises. */
int lrc_grid_connectivity(lrc_ctx* ctx, const double* xs, int32_t nx, const double* ys, int32_t ny, const uint8_t* state,
                          double max_dist, int32_t window, int32_t* free_index, int32_t* row_ptr, int32_t* col,
                          int64_t col_capacity, int64_t* h_counts, void* stream);
/* == _a_star_search (:413-473) over a CSR graph of points, all HOST arrays (the reference's search is host code too):
 *    Euclidean edge costs and heuristic, closed nodes never reopened.  Ties on f are broken by the smaller node index
 *    (the reference's tie-break is CPython's set iteration order), so equal-cost paths may differ in shape, never in
 *    cost.  *h_path_len = 0 when no path exists. */
int lrc_astar(const int32_t* h_row_ptr, const int32_t* h_col, const double* h_pts, int32_t n, int32_t start, int32_t end,
              int32_t* h_path, int32_t path_capacity, int32_t* h_path_len, double* h_cost);

/* ---- measurement ---------------------------------------------------------------------------- */
/* Work counters, accumulated by cast/scan calls while counting is enabled (a separate, slower
 * instantiation of the traversal kernel).  lrc_counters synchronises `stream` first. */
int lrc_set_counting(lrc_ctx* ctx, int enabled);
int lrc_counters(lrc_ctx* ctx, lrc_counters_t* h_out, int reset, void* stream);
/* Number of kernel launches issued by this context since creation (for bench.py's gpu_launches). */
int64_t lrc_launch_count(const lrc_ctx* ctx);
/* Small integers about the context, by name: "scratch_bytes" (grow-only scan/build scratch currently allocated),
 * "bvh_bytes", "node_format", "build_quality", "ploc_iterations" (of the resident tree), the current options "leaf_size",
 * "variant", "tune", "warp_packet", "rays_per_thread", "persistent", "num_sms", and the generation counters
 * "mesh_generation" / "nn_generation" / "collision_generation", which lrc_set_mesh / lrc_nn_index_build /
 * lrc_collision_index_build bump: the context holds ONE index of each kind, so a host object that built one remembers the
 * generation and rebuilds (or refuses) when somebody else has re-targeted the slot since. */
int lrc_get_stat(lrc_ctx* ctx, const char* key, int64_t* h_value);
/* Device time of the traversal kernel (k_trace) and of the ordered compaction (k_scan_counts + k_compact) of the LAST
 * scan call, from CUDA events recorded on the streams those kernels were launched on; summed over the call's pose
 * chunks (*h_launches = number of chunks = k_trace launches).  Requires lrc_set_option(ctx, "kernel_timing", 1)
 * before the scan; synchronises on the recorded events.  This is what bench.py's roofline divides by. */
int lrc_kernel_times(lrc_ctx* ctx, double* h_trace_ms, double* h_compact_ms, int32_t* h_launches);
/* Tuning knobs (defaults are the measured best; none of them changes a result bit):
 *   "variant"       traversal kernel: bit 0 while-while loop, bit 1 256-bit node loads, bit 2 32-register build,
 *                   bit 3 top of the tree in shared memory ("top_levels" 1..8), bit 4 stack in shared memory
 *                   ("stack_levels" 1..48), bit 6 stack entries culled against the best hit at pop time; accepted values
 *                   0..3, 5, 13, 21, 65 (default).  Node format 2 always runs the packed-FMA loop (with / without bit 6).
 *   "leaf_size"     triangles per leaf built by the NEXT lrc_set_mesh, 1..8 (default 2): subtrees with at most that many
 *                   consecutive triangles are tested as one leaf
 *   "node_format"   record format built by the NEXT lrc_set_mesh: 0 = 64 B float boxes, 1 = 32 B 16-bit boxes,
 *                   2 = 64 B paired boxes for the two-wide FP32 FMA of sm_100 (default)
 *   "build_quality" builder of the NEXT lrc_set_mesh: 0 = LBVH (Karras radix tree, default), 1 = PLOC (locally-ordered
 *                   clustering with "ploc_radius" 1..32, default 16): ~12 % fewer node fetches per ray, 2x the build time
 *   "compact_nodes" 1: squeeze the never-read records of collapsed subtrees out of the node array (half the node bytes)
 *   "rays_per_thread" 1 (default), 2, 4: adjacent rays per thread of the scan kernels (format 2 only; measured slower)
 *   "persistent"    1 / 2: persistent warps that fetch 128-ray blocks by ticket instead of one block per 128 rays
 *   "block"         threads per traversal block (32 / 64 / 128); 0 (default) = by the size of the call: 32 up to 2^16 rays,
 *                   64 up to 2^18, 128 beyond
 *   "chunk_rays"    rays per traversal chunk (bounds the 24 B/ray scratch)
 *   "gather_chunks", "gather_ramp", "gather_taper", "push_blocks", "push_mode"   pose chunks / short first chunk / short
 *                   last chunk / exchange blocks / exchange kernel (1 = TMA bulk copies, default; 0 = vector loads and stores)
 *                   of the all-gather
 *   "kernel_timing" record CUDA events around k_trace and the compaction (lrc_kernel_times)
 *   "l2_persist"    percent of the device's maximum persisting-L2 set-aside reserved for an access-policy window over
 *                   the BVH records on every k_trace launch (0 = off, default); "l2_reset" demotes persisting lines now
 * Unknown keys -> LRC_ERR_INVALID. */
int lrc_default_l2_persist(void);
int lrc_set_option(lrc_ctx* ctx, const char* key, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* LRC_H */
