// common.cuh -- context, error plumbing and small device helpers shared by the engine's kernels.
// Target: sm_100a only (B200).  No CPU fallback exists anywhere in this library.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>

#include "../../include/lrc.h"

#define LRC_STACK_DEPTH 64        // per-ray traversal stack entries (checked against the built tree)
#define LRC_MAX_H 4096            // scan lines per single-axis sensor
#define LRC_MAX_GATHER 16         // == lrc_gather's array length
#define LRC_TOP_LEVELS_MAX 8      // top-of-tree table: 2^8 - 1 = 255 node records in heap order (16 KB)

// 32-byte node format (option "node_format" = 1): child boxes as 16-bit cell indices on a per-axis grid over the scene,
// x = o + q * s.  See traverse.cuh (inner_step_q) and bvh_build.cu (write_node_q).
struct NodeQ { float o[3]; float s[3]; };

struct lrc_ctx {
    int device = 0;
    char err[512] = {0};
    int64_t launches = 0;

    // ---- scene (owned) ----
    bool has_mesh = false;
    int64_t mesh_generation = 0;  // bumped by every lrc_set_mesh
    bool has_labels = false;
    int64_t T = 0, V = 0;
    float4* nodes = nullptr;      // num_nodes x 4 float4 (64 B records)
    cudaTextureObject_t nodes_tex = 0;   // the same records behind the texture path (tune bits 3 / 4): its own L1TEX data pipe and write-back
    float4* tris = nullptr;       // T x 3 float4 (48 B records, Morton order): (v0|orig id) (e1|0) (e2|0)
    uint32_t* labels = nullptr;   // T, original triangle order
    int64_t opt_leaf_size = 2;    // triangles per leaf built by the NEXT lrc_set_mesh (1..8); 2 measured best
    int64_t opt_build_quality = 0;   // builder of the NEXT lrc_set_mesh: 0 = LBVH (Karras radix tree), 1 = PLOC (bvh_ploc.cuh)
    int64_t opt_ploc_radius = 16;    // PLOC search radius (clusters before / after in Morton order), 1..32
    int build_quality = 0;        // builder of the tree that is resident now
    int ploc_iterations = 0;
    int root = 0;                 // node record the traversal starts at
    int* h_pin = nullptr;         // small page-locked mailbox (PLOC merge counts)
    int64_t opt_node_format = 2;  // format the NEXT lrc_set_mesh builds: 0 = 64 B float boxes, 1 = 32 B 16-bit boxes, 2 = 64 B paired boxes (packed FMA)
    int64_t opt_compact_nodes = 0;   // 1: keep only the live node records (half the node footprint, +30 % build time, < 1 % traversal time)
    int node_format = 0;          // format of the tree that is resident now
    NodeQ nodeq = {};
    float4* top_table = nullptr;  // (2^LRC_TOP_LEVELS_MAX - 1) x 4 float4: copies of the top nodes in heap order
    int64_t opt_stack_levels = 12;   // traversal-stack entries kept in shared memory by the VARIANT-bit-4 kernel
    int64_t opt_top_levels = 6;   // levels staged in shared memory by the VARIANT-bit-3 traversal kernel
    void* bvh_block = nullptr;    // one allocation [nodes | tris]: a single L2 access-policy window covers both
    size_t bvh_block_bytes = 0;
    size_t labels_cap = 0;        // in elements
    // L2 persistence for the BVH records (option "l2_persist"): k_trace is launched with an access-policy window over
    // bvh_block so that streaming traffic (scratch, outputs, peers' incoming stores) cannot evict the tree
    size_t l2_persist_max = 0, l2_window_max = 0, l2_size = 0;
    int64_t opt_l2_persist = 0;
    size_t bvh_bytes = 0;         // bytes of bvh_block in use
    lrc_bvh_info info = {};
    double root_area = 0.0;

    // ---- grow-only scratch ----
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    cudaEvent_t tables_event = nullptr;    // recorded after the ray tables were (re)built
    cudaEvent_t scratch_event = nullptr;   // recorded after the last scan that used the scratch
    void* scratch2 = nullptr;     // scan bookkeeping (tile status words, ticket, running total)
    size_t scratch2_bytes = 0;
    double* tables = nullptr;     // ray-generation tables
    size_t tables_bytes = 0;
    std::vector<double> tables_key;   // the sensor the resident tables were built for: W, H, mode, fov_up, fov_down, elevations

    // ---- host-buffer entry points (lrc_*_host): internal streams, staging, device outputs ----
    cudaStream_t s_compute = nullptr, s_copy = nullptr, s_count = nullptr;
    void* host_dev = nullptr;          // device-side lrc_out arrays + poses for a whole trajectory
    size_t host_dev_bytes = 0;
    int64_t* h_stage = nullptr;        // page-locked frame-offset staging
    size_t h_stage_elems = 0;
    cudaEvent_t* events = nullptr;     // 2 per chunk
    size_t n_events = 0;
    void* mesh_dev = nullptr;          // device copies of host mesh arrays (lrc_set_mesh_host)
    size_t mesh_dev_bytes = 0;

    // ---- pipelined compaction / multi-GPU gather ----
    cudaStream_t s_aux = nullptr;      // compaction of chunk c runs here while chunk c+1 is traversed
    cudaEvent_t pipe_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    struct {
        int n = 0;
        float* xyz[LRC_MAX_GATHER];
        uint32_t* label[LRC_MAX_GATHER];
        int64_t* frame_offset[LRC_MAX_GATHER];
        int64_t point_base = 0, frame_base = 0, capacity = 0;
        int64_t frame_capacity = 0;
    } gather;
    lrc_gather_wire wire = {};         // compact wire format of the exchange (lrc_set_gather_wire); wire.enabled == 0: off
    int64_t wire_scan = 0;             // scans issued since lrc_set_gather_wire: tags the progress words
    cudaStream_t s_rebuild = nullptr;  // waits for the peers' progress words and rebuilds their points
    cudaEvent_t rebuild_ev = nullptr;
    int* d_wire_err = nullptr;         // set by the waiting kernel when a peer's progress word does not arrive in time
    int64_t opt_gather_chunks = 4;
    int64_t opt_gather_ramp = 1;        // first gather chunk = regular chunk / ramp
    int64_t opt_gather_taper = 1;       // last gather chunk = regular chunk / taper
    int64_t opt_scan_chunks = 1;        // pose chunks of a scan WITHOUT gather targets (compaction of chunk c behind the traversal of c+1)
    int64_t opt_scan_taper = 1;         // ... last chunk = regular chunk / taper
    int64_t opt_push_blocks = 64;       // blocks in total of k_push_tma (push_mode 1) / blocks per target of k_push (push_mode 0; 16 measured best there)
    int64_t opt_push_tile = 16384;      // bytes per stage of k_push_tma (4 stages of shared memory per block)
    bool push_tma_ready = false;        // cudaFuncSetAttribute(k_push_tma, max dynamic shared memory) done on this device
    int64_t opt_push_mode = 1;          // exchange kernel: 1 = k_push_tma (TMA bulk copies, default), 0 = k_push (vector loads / stores)

    // ---- planner support (plan.cu): binned vertex index ----
    bool ci_ready = false;
    int64_t ci_generation = 0;    // bumped by every lrc_collision_index_build: host objects notice that the slot was re-targeted
    int64_t ci_V = 0;
    void* ci_meta = nullptr; size_t ci_meta_bytes = 0;
    void* ci_start = nullptr; size_t ci_start_bytes = 0;
    void* ci_sorted = nullptr; size_t ci_sorted_bytes = 0;
    void* cg_scratch = nullptr; size_t cg_scratch_bytes = 0;     // lrc_grid_connectivity's own scratch
    double ci_ox = 0, ci_oy = 0, ci_cell = 0;
    int ci_nbx = 0, ci_nby = 0;

    // ---- 1-NN label transfer (nn.cu): binned annotated points ----
    bool nn_ready = false;
    int64_t nn_generation = 0;    // bumped by every lrc_nn_index_build
    int64_t nn_n = 0;
    void* nn_meta = nullptr; size_t nn_meta_bytes = 0;
    void* nn_start = nullptr; size_t nn_start_bytes = 0;
    void* nn_sorted = nullptr; size_t nn_sorted_bytes = 0;
    double nn_o[3] = {0, 0, 0}, nn_cell = 0;
    int nn_nb[3] = {0, 0, 0};

    // ---- post-processing (post.cu) ----
    void* post_scratch = nullptr;
    size_t post_scratch_bytes = 0;

    // ---- measurement ----
    unsigned long long* d_counters = nullptr;   // rays, nodes, tris, hits
    int counting = 0;
    // per-kernel CUDA-event timing of the last scan call (option "kernel_timing"): 4 events per pose chunk,
    // [before k_trace, after k_trace] on the caller's stream, [before k_scan_counts, after k_compact] on the
    // stream the compaction runs on
    int opt_kernel_timing = 0;
    std::vector<cudaEvent_t> kt_events;
    size_t kt_used = 0;

    // ---- options ----
    int64_t opt_rays_per_thread = 1;    // scan kernels: adjacent rays per thread (1, 2, 4); > 1 needs node format 2
    int64_t opt_warp_packet = 0;        // scan kernels, format 2: the warp walks the tree together (k_trace_w)
    int64_t opt_tune = 2;               // format-2 kernel: bit 0 prefetch pushed records, bit 1 no block barrier (default), bit 2 streaming scratch stores
    int64_t opt_persistent = 0;         // 1: persistent warps over 32-ray tiles (VARIANT bit 8) instead of one block per 128 rays
    int num_sms = 148;
    int64_t opt_block = 0;              // threads per traversal block (32 / 64 / 128); 0 = by the size of the call
    int cur_block = 128;                // block size of the scan call being enqueued
    int64_t opt_chunk_rays = 1 << 26;   // rays per traversal/epilogue chunk (bounds scratch: 16 B per ray)
    int64_t opt_variant = 65;           // traversal kernel variant: while-while loop + stack entries culled at pop time (bit 6);
                                        // with node format 2 the packed-FMA loop (bit 7) is selected automatically
};

extern char g_lrc_global_err[512];

static inline int lrc_fail(lrc_ctx* ctx, int code, const char* fmt, const char* a = "", const char* b = "")
{
    char* dst = ctx ? ctx->err : g_lrc_global_err;
    snprintf(dst, 512, fmt, a, b);
    return code;
}

#define LRC_CUDA(ctx, call)                                                                      \
    do {                                                                                         \
        cudaError_t e__ = (call);                                                                \
        if (e__ != cudaSuccess) return lrc_fail(ctx, LRC_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
    } while (0)

#define LRC_CHECK_LAUNCH(ctx, name)                                                              \
    do {                                                                                         \
        (ctx)->launches++;                                                                       \
        cudaError_t e__ = cudaGetLastError();                                                    \
        if (e__ != cudaSuccess) return lrc_fail(ctx, LRC_ERR_CUDA, "launch %s: %s", name, cudaGetErrorString(e__)); \
    } while (0)

static inline int lrc_grow(lrc_ctx* ctx, void** p, size_t* cap_bytes, size_t need_bytes)
{
    if (*cap_bytes >= need_bytes && *p) return LRC_OK;
    if (*p) { LRC_CUDA(ctx, cudaFree(*p)); *p = nullptr; *cap_bytes = 0; }
    size_t n = need_bytes < 256 ? 256 : need_bytes;
    LRC_CUDA(ctx, cudaMalloc(p, n));
    *cap_bytes = n;
    return LRC_OK;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers -----------------------------------------------------------------------------
#define LRC_INF __int_as_float(0x7f800000)

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

// order-preserving float <-> uint mapping for atomicMin/atomicMax on floats
__device__ __forceinline__ unsigned f2ord(float f)
{
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// Philox4x32-10 (Salmon et al. SC'11).  key = seed, counter = (ray, pose_lo, pose_hi, stream).
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0);
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return c;
}
__device__ __forceinline__ double u01(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }
