// bvh_build.cu -- GPU construction of the linear BVH that replaces
//     o3d.t.geometry.RaycastingScene() + add_triangles(...)   (reference raycast_engine_cpu.py:46-47)
//
// Pipeline (all on the caller's stream, one synchronisation at the end to read back tree depth):
//   k_vertex_bounds, k_check_indices   scene AABB (warp shuffles + ordered-int atomics), index validation
//   k_morton         Morton code of each triangle's box centre (cubic cells, 16 bits/axis kept of 21)
//   radix sort       hand-written LSD sort of (u64 key, u32 triangle id), 8 bits x 6 passes (48 key bits):
//                    k_rs_hist -> k_rs_scan -> k_rs_scatter (stable multi-split by warp match)
//   k_leaf_init      48 B triangle records (v0|id, e1, e2) in Morton order + padded leaf boxes
//   k_hierarchy      Karras 2012 binary radix tree over the sorted keys (ties broken by position)
//   k_refit          bottom-up AABB union with atomic arrival flags; the second arriver writes the
//                    64 B node record (both child boxes + child links), tree height and SAH sum
//
// HBM layout produced:
//   nodes: (T-1) x 4 float4   child boxes as centre c / half-extent h (h rounded up):
//                             n0=(c0.xyz, h0.x) n1=(h0.yz, c1.xy) n2=(c1.z, h1.xyz)
//                             n3=(link0, link1, 0, 0) as int bits; link >= 0 internal, < 0 leaf ~slot
//   tris : T x 3 float4       (v0.xyz, orig id) (e1.xyz, 0) (e2.xyz, 0), slot = Morton rank
#include "common.cuh"

namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_CHUNKS = 16;
constexpr int RS_TILE = RS_THREADS * RS_CHUNKS;

struct BuildMeta {            // lives in device memory during a build
    unsigned lo[3], hi[3];    // ordered-uint encoded scene bounds
    int bad_index;            // set when a triangle references a vertex outside [0, V)
    int height;               // tree height (edges on the longest root->leaf path)
    float root_lo[3], root_hi[3];
    int root;                 // index of the root's node record (0 for the radix tree; in-order index for PLOC)
};

__global__ void k_meta_init(BuildMeta* m)
{
    for (int k = 0; k < 3; ++k) { m->lo[k] = 0xffffffffu; m->hi[k] = 0u; }
    m->bad_index = 0;
    m->height = 0;
    m->root = 0;
}

// scene AABB over the vertex array (coalesced; a vertex no triangle references still counts -- harmless for the
// Morton normalisation and the pad) ...
__global__ void k_vertex_bounds(const float* __restrict__ verts, int64_t V, BuildMeta* meta)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {LRC_INF, LRC_INF, LRC_INF}, hi[3] = {-LRC_INF, -LRC_INF, -LRC_INF};
    if (i < V) {
#pragma unroll
        for (int k = 0; k < 3; ++k) lo[k] = hi[k] = verts[3 * i + k];
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float l = lo[k], h = hi[k];
        // a NaN coordinate must not vanish inside fminf/fmaxf: turn it into +inf so the finiteness check fires
        if (l != l) { l = LRC_INF; h = LRC_INF; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l = fminf(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = fmaxf(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if (lane_id() == 0 && l <= h) {
            atomicMin(&meta->lo[k], f2ord(l));
            atomicMax(&meta->hi[k], f2ord(h));
        }
    }
}

// ... and index validation over the 3T indices (coalesced)
__global__ void k_check_indices(const int32_t* __restrict__ tris, int64_t n3, int64_t V, BuildMeta* meta)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool bad = false;
    if (i < n3) {
        int64_t v = tris[i];
        bad = v < 0 || v >= V;
    }
    if (__any_sync(0xffffffffu, bad) && lane_id() == 0) atomicExch(&meta->bad_index, 1);
}

__device__ __forceinline__ uint64_t expand21(uint32_t v)
{
    uint64_t x = v & 0x1fffffu;
    x = (x | x << 32) & 0x001f00000000ffffull;
    x = (x | x << 16) & 0x001f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float* __restrict__ verts, const int32_t* __restrict__ tris, int64_t T,
                         const BuildMeta* __restrict__ meta, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    float slo[3], ext = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        slo[k] = ord2f(meta->lo[k]);
        ext = fmaxf(ext, ord2f(meta->hi[k]) - slo[k]);
    }
    float scale = ext > 0.f ? 2097152.0f / ext : 0.f;   // 2^21 cubic cells along the longest axis
    uint32_t q[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float a = verts[3 * (int64_t)tris[3 * i + 0] + k];
        float b = verts[3 * (int64_t)tris[3 * i + 1] + k];
        float c = verts[3 * (int64_t)tris[3 * i + 2] + k];
        float cen = 0.5f * (fminf(a, fminf(b, c)) + fmaxf(a, fmaxf(b, c)));
        float f = (cen - slo[k]) * scale;
        f = fminf(fmaxf(f, 0.f), 2097151.0f);
        q[k] = (uint32_t)f;
    }
    // 16 bits per axis are kept (cells of 2^-16 of the scene, 0.3 mm for a 20 m room): the low 15 bits of the 63-bit code
    // are dropped so that the sort needs 6 instead of 8 byte passes; equal keys are ordered by position (k_hierarchy)
    keys[i] = ((expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2])) & ~0x7fffull;
    vals[i] = (uint32_t)i;
}

// ---- LSD radix sort -----------------------------------------------------------------------------
// Tile = 4096 keys per block; warp w owns 512 consecutive keys and walks them 32 at a time, so the
// (warp, chunk, lane) order IS the input order and ranks computed from it are stable.
__device__ __forceinline__ void rs_count_tile(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                              int64_t tile_start, unsigned (*hist)[256])
{
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&hist[0][0])[i] = 0u;
    __syncthreads();
    const int64_t base = tile_start + (int64_t)w * (32 * RS_CHUNKS);
    for (int c = 0; c < RS_CHUNKS; ++c) {
        int64_t idx = base + c * 32 + lane;
        bool valid = idx < n;
        unsigned digit = valid ? (unsigned)((keys[idx] >> shift) & 255ull) : 256u;
        unsigned m = __match_any_sync(0xffffffffu, digit);
        if (valid && lane == (unsigned)(__ffs(m) - 1)) hist[w][digit] += __popc(m);
        __syncwarp();
    }
    __syncthreads();
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                        unsigned* __restrict__ block_hist, int nb)
{
    __shared__ unsigned hist[RS_WARPS][256];
    rs_count_tile(keys, n, shift, (int64_t)blockIdx.x * RS_TILE, hist);
    unsigned s = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) s += hist[w][threadIdx.x];
    block_hist[(size_t)threadIdx.x * nb + blockIdx.x] = s;   // digit-major
}

__global__ void __launch_bounds__(1024) k_rs_scan(unsigned* __restrict__ a, int64_t M)
{
    // exclusive scan of M counters by one block, 16 per thread per trip as 4 x uint4 (M = 256 * number of sort tiles,
    // the array is 256 B aligned)
    constexpr int ITEMS = 16;
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (int64_t base = 0; base < M; base += 1024 * ITEMS) {
        const int64_t idx = base + (int64_t)ITEMS * threadIdx.x;
        unsigned v[ITEMS];
        const bool full = idx + ITEMS <= M;
        if (full) {
#pragma unroll
            for (int k = 0; k < ITEMS / 4; ++k) {
                const uint4 q = reinterpret_cast<const uint4*>(a + idx)[k];
                v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) v[k] = idx + k < M ? a[idx + k] : 0u;
        }
        unsigned mine = 0u;
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) mine += v[k];
        unsigned incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[w] = incl;
        __syncthreads();
        if (w == 0) {
            unsigned sv = warp_sums[lane], si = sv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                unsigned t = __shfl_up_sync(0xffffffffu, si, o);
                if (lane >= o) si += t;
            }
            warp_sums[lane] = si - sv;
        }
        __syncthreads();
        unsigned run = incl - mine + warp_sums[w] + carry;
        if (full) {
#pragma unroll
            for (int k = 0; k < ITEMS / 4; ++k) {
                uint4 q;
                q.x = run; run += v[4 * k];
                q.y = run; run += v[4 * k + 1];
                q.z = run; run += v[4 * k + 2];
                q.w = run; run += v[4 * k + 3];
                reinterpret_cast<uint4*>(a + idx)[k] = q;
            }
        } else {
#pragma unroll
            for (int k = 0; k < ITEMS; ++k) {
                if (idx + k < M) a[idx + k] = run;
                run += v[k];
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = run;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(RS_THREADS)
k_rs_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint64_t* __restrict__ keys_out,
             uint32_t* __restrict__ vals_out, int64_t n, int shift, const unsigned* __restrict__ scanned, int nb)
{
    __shared__ unsigned hist[RS_WARPS][256];
    const int64_t tile_start = (int64_t)blockIdx.x * RS_TILE;
    rs_count_tile(keys_in, n, shift, tile_start, hist);
    {   // per-digit: turn warp counts into global base offsets
        unsigned run = scanned[(size_t)threadIdx.x * nb + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            unsigned c = hist[w][threadIdx.x];
            hist[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t base = tile_start + (int64_t)w * (32 * RS_CHUNKS);
    for (int c = 0; c < RS_CHUNKS; ++c) {
        int64_t idx = base + c * 32 + lane;
        bool valid = idx < n;
        uint64_t key = valid ? keys_in[idx] : 0ull;
        unsigned digit = valid ? (unsigned)((key >> shift) & 255ull) : 256u;
        unsigned m = __match_any_sync(0xffffffffu, digit);
        unsigned rank = __popc(m & ((1u << lane) - 1u));
        unsigned dst = 0;
        if (valid) dst = hist[w][digit] + rank;
        __syncwarp();
        if (valid && lane == (unsigned)(__ffs(m) - 1)) hist[w][digit] += __popc(m);
        __syncwarp();
        if (valid) {
            keys_out[dst] = key;
            vals_out[dst] = vals_in[idx];
        }
    }
}

// ---- leaves -------------------------------------------------------------------------------------
__global__ void k_leaf_init(const float* __restrict__ verts, const int32_t* __restrict__ tris, int64_t T,
                            const uint32_t* __restrict__ sorted_ids, float pad, float4* __restrict__ tri_rec,
                            float4* __restrict__ leaf_lo, float4* __restrict__ leaf_hi)
{   // tri_rec == nullptr: boxes only (the PLOC builder writes the records later, in depth-first leaf order)
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= T) return;
    uint32_t id = sorted_ids[i];
    float a[3], b[3], c[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a[k] = verts[3 * (int64_t)tris[3 * (int64_t)id + 0] + k];
        b[k] = verts[3 * (int64_t)tris[3 * (int64_t)id + 1] + k];
        c[k] = verts[3 * (int64_t)tris[3 * (int64_t)id + 2] + k];
    }
    // e1 = v1 - v0, e2 = v2 - v0: one IEEE subtraction each -- part of the intersection spec
    if (tri_rec) {
        tri_rec[3 * i + 0] = make_float4(a[0], a[1], a[2], __uint_as_float(id));
        tri_rec[3 * i + 1] = make_float4(__fsub_rn(b[0], a[0]), __fsub_rn(b[1], a[1]), __fsub_rn(b[2], a[2]), 0.f);
        tri_rec[3 * i + 2] = make_float4(__fsub_rn(c[0], a[0]), __fsub_rn(c[1], a[1]), __fsub_rn(c[2], a[2]), 0.f);
    }
    leaf_lo[i] = make_float4(fminf(a[0], fminf(b[0], c[0])) - pad, fminf(a[1], fminf(b[1], c[1])) - pad,
                             fminf(a[2], fminf(b[2], c[2])) - pad, 0.f);
    leaf_hi[i] = make_float4(fmaxf(a[0], fmaxf(b[0], c[0])) + pad, fmaxf(a[1], fmaxf(b[1], c[1])) + pad,
                             fmaxf(a[2], fmaxf(b[2], c[2])) + pad, 0.f);
}

// ---- Karras 2012 --------------------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz(i ^ j);
    return __clzll((long long)(a ^ b));
}

__global__ void k_hierarchy(const uint64_t* __restrict__ keys, int n, int2* __restrict__ children,
                            int* __restrict__ parent_node, int* __restrict__ parent_leaf, int2* __restrict__ range)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    int gamma = i + s * d + min(d, 0);
    int lo = min(i, j), hi = max(i, j);
    int left = (lo == gamma) ? ~gamma : gamma;
    int right = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    range[i] = make_int2(lo, hi - lo + 1);          // the Morton-contiguous leaf slots [lo, hi] under this node
    if (left < 0) parent_leaf[~left] = i; else parent_node[left] = i;
    if (right < 0) parent_leaf[~right] = i; else parent_node[right] = i;
    if (i == 0) parent_node[0] = -1;
}

__device__ __forceinline__ void load_box(int link, const float4* leaf_lo, const float4* leaf_hi, const float4* node_lo,
                                         const float4* node_hi, float4& lo, float4& hi)
{
    // __ldcg: these boxes may have been written by another SM moments ago -- bypass L1
    if (link < 0) { lo = __ldcg(&leaf_lo[~link]); hi = __ldcg(&leaf_hi[~link]); }
    else { lo = __ldcg(&node_lo[link]); hi = __ldcg(&node_hi[link]); }
}

// centre / half-extent form of a box; the half-extent is rounded UP so that [c-h, c+h] contains [lo, hi]
__device__ __forceinline__ void to_centre_half(const float4 lo, const float4 hi, float c[3], float h[3])
{
    const float l[3] = {lo.x, lo.y, lo.z}, u[3] = {hi.x, hi.y, hi.z};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        c[k] = 0.5f * (l[k] + u[k]);
        h[k] = fmaxf(__fsub_ru(u[k], c[k]), __fsub_ru(c[k], l[k]));
    }
}

__device__ __forceinline__ void write_node(float4* __restrict__ rec, const float4 l0, const float4 h0, const float4 l1,
                                           const float4 h1, int link0, int link1)
{
    float c0[3], e0[3], c1[3], e1[3];
    to_centre_half(l0, h0, c0, e0);
    to_centre_half(l1, h1, c1, e1);
    rec[0] = make_float4(c0[0], c0[1], c0[2], e0[0]);
    rec[1] = make_float4(e0[1], e0[2], c1[0], c1[1]);
    rec[2] = make_float4(c1[2], e1[0], e1[1], e1[2]);
    rec[3] = make_float4(__int_as_float(link0), __int_as_float(link1), 0.f, 0.f);
}

// Format 2, "paired" 64 B record for the packed-FMA traversal (fma.rn.f32x2 -> FFMA2 on sm_100): the values that meet the
// same pair of ray constants sit in adjacent, 8-byte aligned words
//     n0 = (c0.x, c0.y, c1.x, c1.y)   n1 = (h0.x, h0.y, h1.x, h1.y)   n2 = (c0.z, c1.z, h0.z, h1.z)   n3 = (link0, link1, 0, 0)
__device__ __forceinline__ void write_node_p(float4* __restrict__ rec, const float4 l0, const float4 h0, const float4 l1,
                                             const float4 h1, int link0, int link1)
{
    float c0[3], e0[3], c1[3], e1[3];
    to_centre_half(l0, h0, c0, e0);
    to_centre_half(l1, h1, c1, e1);
    rec[0] = make_float4(c0[0], c0[1], c1[0], c1[1]);
    rec[1] = make_float4(e0[0], e0[1], e1[0], e1[1]);
    rec[2] = make_float4(c0[2], c1[2], e0[2], e1[2]);
    rec[3] = make_float4(__int_as_float(link0), __int_as_float(link1), 0.f, 0.f);
}

// 32-byte record: (x0 y0 z0 link0)(x1 y1 z1 link1), each axis word = qlo | qhi << 16 with x = o + q * s.
// Conservative by construction: lo is rounded down and hi up to the cell grid and both are moved out by one more cell,
// which covers the rounding of this division (< 0.01 cell) and of the traversal's decode (< 0.6 cell, traverse.cuh).
__device__ __forceinline__ unsigned quant_axis(float lo, float hi, float o, float s)
{
    int a = (int)floorf((lo - o) / s) - 1, b = (int)ceilf((hi - o) / s) + 1;
    a = min(max(a, 0), 65535);
    b = min(max(b, 0), 65535);
    return (unsigned)a | ((unsigned)b << 16);
}

__device__ __forceinline__ void write_node_q(float4* __restrict__ rec, const NodeQ& q, const float4 l0, const float4 h0,
                                             const float4 l1, const float4 h1, int link0, int link1)
{
    rec[0] = make_float4(__uint_as_float(quant_axis(l0.x, h0.x, q.o[0], q.s[0])), __uint_as_float(quant_axis(l0.y, h0.y, q.o[1], q.s[1])),
                         __uint_as_float(quant_axis(l0.z, h0.z, q.o[2], q.s[2])), __int_as_float(link0));
    rec[1] = make_float4(__uint_as_float(quant_axis(l1.x, h1.x, q.o[0], q.s[0])), __uint_as_float(quant_axis(l1.y, h1.y, q.o[1], q.s[1])),
                         __uint_as_float(quant_axis(l1.z, h1.z, q.o[2], q.s[2])), __int_as_float(link1));
}

// one node record in the format of the tree being built (0: centre/half float4 x 4, 1: 16-bit boxes, 2: paired)
__device__ __forceinline__ void emit_node(float4* __restrict__ nodes_out, int64_t idx, int format, const NodeQ& nq, const float4 l0,
                                          const float4 h0, const float4 l1, const float4 h1, int link0, int link1)
{
    if (format == 0) write_node(nodes_out + 4 * idx, l0, h0, l1, h1, link0, link1);
    else if (format == 2) write_node_p(nodes_out + 4 * idx, l0, h0, l1, h1, link0, link1);
    else write_node_q(nodes_out + 2 * idx, nq, l0, h0, l1, h1, link0, link1);
}

__global__ void k_refit(int n, const int* __restrict__ parent_leaf, const int* __restrict__ parent_node,
                        const int2* __restrict__ children, const float4* leaf_lo, const float4* leaf_hi, float4* node_lo,
                        float4* node_hi, int* flags, float4* __restrict__ nodes_out, BuildMeta* meta, int format, NodeQ nq,
                        const int2* __restrict__ range, int leaf_max)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cur = parent_leaf[i];
    while (cur >= 0) {
        if (atomicAdd(&flags[cur], 1) == 0) return;   // first arrival: the sibling subtree is not finished yet
        int2 ch = children[cur];
        float4 l0, h0, l1, h1;
        load_box(ch.x, leaf_lo, leaf_hi, node_lo, node_hi, l0, h0);
        load_box(ch.y, leaf_lo, leaf_hi, node_lo, node_hi, l1, h1);
        float height = 1.f + fmaxf(l0.w, l1.w);        // .w of a box's lo carries the subtree height
        float4 lo = make_float4(fminf(l0.x, l1.x), fminf(l0.y, l1.y), fminf(l0.z, l1.z), height);
        float4 hi = make_float4(fmaxf(h0.x, h1.x), fmaxf(h0.y, h1.y), fmaxf(h0.z, h1.z), 0.f);
        node_lo[cur] = lo;
        node_hi[cur] = hi;
        // a child subtree with at most leaf_max triangles becomes ONE leaf over its contiguous slots:
        // link = ~(first slot | (count - 1) << 28); a single-triangle leaf keeps the plain ~slot form
        int k0 = ch.x, k1 = ch.y;
        if (k0 >= 0 && range[k0].y <= leaf_max) k0 = ~(range[k0].x | ((range[k0].y - 1) << 28));
        if (k1 >= 0 && range[k1].y <= leaf_max) k1 = ~(range[k1].x | ((range[k1].y - 1) << 28));
        emit_node(nodes_out, cur, format, nq, l0, h0, l1, h1, k0, k1);
        int up = parent_node[cur];
        if (up < 0) {
            meta->height = (int)height;
            meta->root_lo[0] = lo.x; meta->root_lo[1] = lo.y; meta->root_lo[2] = lo.z;
            meta->root_hi[0] = hi.x; meta->root_hi[1] = hi.y; meta->root_hi[2] = hi.z;
        }
        __threadfence();
        cur = up;
    }
}

// ---- live-node compaction ---------------------------------------------------------------------------------------------
// A subtree of <= leaf_size triangles is ONE leaf link in its parent, so the records of its inner nodes are never read --
// with two-triangle leaves that is about half of the T-1 records, interleaved with the live ones: every 128 B line of the
// bottom levels would carry one live record on average.  The builders therefore write their records to scratch, and
// k_compact_nodes moves the live ones (root + nodes with more than leaf_size triangles) to consecutive slots in the same
// order (siblings stay adjacent, subtrees stay contiguous) and rewrites the inner links: half the node footprint in L1 / L2
// / HBM for the same tree.
__global__ void k_live_flags_lbvh(int n_nodes, const int2* __restrict__ range, int leaf_max, unsigned* __restrict__ live)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_nodes) live[i] = (i == 0 || range[i].y > leaf_max) ? 1u : 0u;
}

__global__ void k_compact_nodes(int n_nodes, const float4* __restrict__ src, const unsigned* __restrict__ new_index,
                                const unsigned* __restrict__ live, float4* __restrict__ dst, int format)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes || !live[i]) return;
    const int64_t d = new_index[i];
    if (format == 1) {
        float4 a = src[2 * (int64_t)i], b = src[2 * (int64_t)i + 1];
        const int l0 = __float_as_int(a.w), l1 = __float_as_int(b.w);
        if (l0 >= 0) a.w = __int_as_float((int)new_index[l0]);
        if (l1 >= 0) b.w = __int_as_float((int)new_index[l1]);
        dst[2 * d] = a; dst[2 * d + 1] = b;
    } else {
        const float4* r = src + 4 * (int64_t)i;
        float4 k = r[3];
        const int l0 = __float_as_int(k.x), l1 = __float_as_int(k.y);
        if (l0 >= 0) k.x = __int_as_float((int)new_index[l0]);
        if (l1 >= 0) k.y = __int_as_float((int)new_index[l1]);
        dst[4 * d] = r[0]; dst[4 * d + 1] = r[1]; dst[4 * d + 2] = r[2]; dst[4 * d + 3] = k;
    }
}

// T == 1: a root record whose two links both point at the only leaf (testing it twice is harmless)
__global__ void k_single_leaf_root(const float4* leaf_lo, const float4* leaf_hi, float4* nodes_out, BuildMeta* meta, int format,
                                   NodeQ nq)
{
    float4 l = leaf_lo[0], h = leaf_hi[0];
    emit_node(nodes_out, 0, format, nq, l, h, l, h, ~0, ~0);
    meta->height = 1;
    meta->root_lo[0] = l.x; meta->root_lo[1] = l.y; meta->root_lo[2] = l.z;
    meta->root_hi[0] = h.x; meta->root_hi[1] = h.y; meta->root_hi[2] = h.z;
}

// Heap-ordered copy of the top K node records (entry h: children at 2h+1, 2h+2) for the traversal variant that stages
// the top of the tree in shared memory.  One block; levels are resolved one after the other.
__global__ void __launch_bounds__(256) k_top_table(const float4* __restrict__ nodes, float4* __restrict__ top, int K, int root)
{
    __shared__ int gid[256];
    for (int h = threadIdx.x; h < 256; h += blockDim.x) gid[h] = h == 0 ? root : -1;
    __syncthreads();
    for (int s = 1; s < K; s = 2 * s + 1) {                 // level [s, 2s+1)
        for (int h = s + threadIdx.x; h < min(2 * s + 1, K); h += blockDim.x) {
            const int p = (h - 1) >> 1;
            if (gid[p] >= 0) {
                const float4 n3 = nodes[4 * (int64_t)gid[p] + 3];
                const int link = __float_as_int((h & 1) ? n3.x : n3.y);
                gid[h] = link >= 0 ? link : -1;             // a leaf child has no record
            }
        }
        __syncthreads();
    }
    for (int h = threadIdx.x; h < K; h += blockDim.x)
        for (int k = 0; k < 4; ++k)
            top[4 * h + k] = gid[h] >= 0 ? nodes[4 * (int64_t)gid[h] + k] : make_float4(0.f, 0.f, 0.f, 0.f);
}

// SAH diagnostic: sum of the surface areas of every child box stored in the node records
__global__ void k_sah_sum(const float4* __restrict__ nodes, int64_t n_nodes, double* out, int format, NodeQ nq)
{
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_nodes; i += (int64_t)gridDim.x * blockDim.x) {
        if (format == 1) {
            for (int c = 0; c < 2; ++c) {
                const float4 r = nodes[2 * i + c];
                const unsigned w[3] = {__float_as_uint(r.x), __float_as_uint(r.y), __float_as_uint(r.z)};
                double e[3];
                for (int k = 0; k < 3; ++k) e[k] = (double)((w[k] >> 16) - (w[k] & 0xffffu)) * nq.s[k];
                acc += 2.0 * (e[0] * e[1] + e[1] * e[2] + e[2] * e[0]);
            }
            continue;
        }
        const float4 n0 = nodes[4 * i], n1 = nodes[4 * i + 1], n2 = nodes[4 * i + 2];
        if (format == 2) {
            acc += 8.0 * ((double)n1.x * n1.y + (double)n1.y * n2.z + (double)n2.z * n1.x);   // child 0: h = (n1.x, n1.y, n2.z)
            acc += 8.0 * ((double)n1.z * n1.w + (double)n1.w * n2.w + (double)n2.w * n1.z);   // child 1: h = (n1.z, n1.w, n2.w)
            continue;
        }
        acc += 8.0 * ((double)n0.w * n1.x + (double)n1.x * n1.y + (double)n1.y * n0.w);   // child 0: h = (n0.w, n1.x, n1.y)
        acc += 8.0 * ((double)n2.y * n2.z + (double)n2.z * n2.w + (double)n2.w * n2.y);   // child 1: h = (n2.y, n2.z, n2.w)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane_id() == 0) atomicAdd(out, acc);
}

}  // namespace

#include "bvh_ploc.cuh"

// -------------------------------------------------------------------------------------------------
extern "C" int lrc_set_mesh(lrc_ctx* ctx, const float* verts, int64_t V, const int32_t* tris, int64_t T,
                            const uint32_t* tri_label, void* stream_)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_set_mesh: ctx is NULL");
    if (V < 0 || T < 0 || T >= (int64_t)1 << 28) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_mesh: bad V/T (T must be < 2^28)");
    if ((T > 0) && (!verts || !tris)) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_mesh: verts/tris is NULL");
    cudaStream_t stream = (cudaStream_t)stream_;
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->has_mesh = false;
    ctx->mesh_generation++;
    ctx->T = T;
    ctx->V = V;
    ctx->has_labels = tri_label != nullptr;
    memset(&ctx->info, 0, sizeof ctx->info);
    if (T == 0) {   // an empty scene is legal: every ray misses
        ctx->has_mesh = true;
        return LRC_OK;
    }
    // nodes, triangles, labels and the build scratch (carved out of ctx->scratch) are all read or written by scans:
    // order this build after the last scan, whatever stream that ran on
    if (ctx->scratch_event) LRC_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->scratch_event, 0));
    const int64_t n_nodes = T > 1 ? T - 1 : 1;
    const int format = (int)ctx->opt_node_format;
    {
        const size_t nodes_bytes = align_up(sizeof(float4) * (format == 1 ? 2 : 4) * (size_t)n_nodes, 256);
        const size_t tris_bytes = align_up(sizeof(float4) * 3 * (size_t)T, 256);
        int rcb = lrc_grow(ctx, &ctx->bvh_block, &ctx->bvh_block_bytes, nodes_bytes + tris_bytes);
        if (rcb) return rcb;
        ctx->nodes = (float4*)ctx->bvh_block;
        ctx->tris = (float4*)((char*)ctx->bvh_block + nodes_bytes);
        ctx->bvh_bytes = nodes_bytes + tris_bytes;
        // a linear float4 texture over the node records (format 0 / 2: 4 texels per record) for the kernels that split
        // their node loads between the LSU and the texture pipe; beyond the linear-texture limit those kernels are not used
        if (ctx->nodes_tex) { cudaDestroyTextureObject(ctx->nodes_tex); ctx->nodes_tex = 0; }
        if (nodes_bytes / sizeof(float4) <= ((size_t)1 << 27)) {
            cudaResourceDesc rd;
            memset(&rd, 0, sizeof rd);
            rd.resType = cudaResourceTypeLinear;
            rd.res.linear.devPtr = ctx->nodes;
            rd.res.linear.desc = cudaCreateChannelDesc<float4>();
            rd.res.linear.sizeInBytes = nodes_bytes;
            cudaTextureDesc td;
            memset(&td, 0, sizeof td);
            td.readMode = cudaReadModeElementType;
            if (cudaCreateTextureObject(&ctx->nodes_tex, &rd, &td, nullptr) != cudaSuccess) { ctx->nodes_tex = 0; cudaGetLastError(); }
        }
    }
    if ((size_t)T > ctx->labels_cap) {
        if (ctx->labels) LRC_CUDA(ctx, cudaFree(ctx->labels));
        ctx->labels = nullptr; ctx->labels_cap = 0;
        LRC_CUDA(ctx, cudaMalloc((void**)&ctx->labels, sizeof(uint32_t) * T));
        ctx->labels_cap = (size_t)T;
    }
    if (tri_label) LRC_CUDA(ctx, cudaMemcpyAsync(ctx->labels, tri_label, sizeof(uint32_t) * T, cudaMemcpyDeviceToDevice, stream));
    else LRC_CUDA(ctx, cudaMemsetAsync(ctx->labels, 0, sizeof(uint32_t) * T, stream));

    const int quality = (T > 2) ? (int)ctx->opt_build_quality : 0;
    const bool compact = ctx->opt_compact_nodes != 0;

    // ---- carve the build scratch ----
    const int nb = (int)((T + RS_TILE - 1) / RS_TILE);
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    size_t o_meta = carve(sizeof(BuildMeta));
    size_t o_k0 = carve(sizeof(uint64_t) * T), o_k1 = carve(sizeof(uint64_t) * T);
    size_t o_v0 = carve(sizeof(uint32_t) * T), o_v1 = carve(sizeof(uint32_t) * T);
    size_t o_hist = carve(sizeof(unsigned) * 256 * (size_t)nb);
    size_t o_llo = carve(sizeof(float4) * T), o_lhi = carve(sizeof(float4) * T);
    size_t o_nlo = carve(sizeof(float4) * n_nodes), o_nhi = carve(sizeof(float4) * n_nodes);
    size_t o_pl = carve(sizeof(int) * T), o_pn = carve(sizeof(int) * n_nodes);
    size_t o_ch = carve(sizeof(int2) * n_nodes), o_fl = carve(sizeof(int) * n_nodes);
    size_t o_rg = carve(sizeof(int2) * n_nodes);
    size_t o_tmpn = carve(compact ? sizeof(float4) * (format == 1 ? 2 : 4) * (size_t)n_nodes : 16);       // records before the live-node compaction
    size_t o_live = carve(sizeof(unsigned) * (size_t)n_nodes), o_newi = carve(sizeof(unsigned) * ((size_t)n_nodes + 16));
    // PLOC only: two cluster buffers, per-iteration bookkeeping, subtree positions (o_ch / o_rg hold left|right and
    // left-count|start; the key buffer k1 is free after the sort)
    size_t o_cid0 = 0, o_cid1 = 0, o_clo0 = 0, o_clo1 = 0, o_chi0 = 0, o_chi1 = 0, o_nn = 0, o_role = 0, o_bsum = 0, o_sl = 0;
    if (quality) {
        o_cid0 = carve(sizeof(int) * T); o_cid1 = carve(sizeof(int) * T);
        o_clo0 = carve(sizeof(float4) * T); o_clo1 = carve(sizeof(float4) * T);
        o_chi0 = carve(sizeof(float4) * T); o_chi1 = carve(sizeof(float4) * T);
        o_nn = carve(sizeof(int) * T); o_role = carve((size_t)T);
        o_bsum = carve(sizeof(uint2) * (size_t)((T + PL_THREADS - 1) / PL_THREADS));
        o_sl = carve(sizeof(int) * T);
    }
    int rc = lrc_grow(ctx, &ctx->scratch, &ctx->scratch_bytes, off);
    if (rc) return rc;
    char* base = (char*)ctx->scratch;
    BuildMeta* meta = (BuildMeta*)(base + o_meta);
    uint64_t* k0 = (uint64_t*)(base + o_k0); uint64_t* k1 = (uint64_t*)(base + o_k1);
    uint32_t* v0 = (uint32_t*)(base + o_v0); uint32_t* v1 = (uint32_t*)(base + o_v1);
    unsigned* hist = (unsigned*)(base + o_hist);
    float4* leaf_lo = (float4*)(base + o_llo); float4* leaf_hi = (float4*)(base + o_lhi);
    float4* node_lo = (float4*)(base + o_nlo); float4* node_hi = (float4*)(base + o_nhi);
    int* parent_leaf = (int*)(base + o_pl); int* parent_node = (int*)(base + o_pn);
    int2* children = (int2*)(base + o_ch); int* flags = (int*)(base + o_fl);
    int2* range = (int2*)(base + o_rg);
    float4* tmp_nodes = compact ? (float4*)(base + o_tmpn) : ctx->nodes;      // where the builders write their records
    unsigned* live = (unsigned*)(base + o_live);
    unsigned* new_index = (unsigned*)(base + o_newi);

    const int TB = 256;
    const unsigned gT = (unsigned)((T + TB - 1) / TB);
    k_meta_init<<<1, 1, 0, stream>>>(meta);
    LRC_CHECK_LAUNCH(ctx, "k_meta_init");
    k_vertex_bounds<<<(unsigned)((V + TB - 1) / TB), TB, 0, stream>>>(verts, V, meta);
    LRC_CHECK_LAUNCH(ctx, "k_vertex_bounds");
    k_check_indices<<<(unsigned)((3 * T + TB - 1) / TB), TB, 0, stream>>>(tris, 3 * T, V, meta);
    LRC_CHECK_LAUNCH(ctx, "k_check_indices");

    BuildMeta hm;
    LRC_CUDA(ctx, cudaMemcpyAsync(&hm, meta, sizeof hm, cudaMemcpyDeviceToHost, stream));
    LRC_CUDA(ctx, cudaStreamSynchronize(stream));
    if (hm.bad_index) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_mesh: a triangle references a vertex outside [0, V)");
    float slo[3], shi[3], ext = 0.f, amax = 0.f;
    for (int k = 0; k < 3; ++k) {
        unsigned ul = hm.lo[k], uh = hm.hi[k];
        uint32_t bl = (ul & 0x80000000u) ? (ul & 0x7fffffffu) : ~ul;
        uint32_t bh = (uh & 0x80000000u) ? (uh & 0x7fffffffu) : ~uh;
        memcpy(&slo[k], &bl, 4);
        memcpy(&shi[k], &bh, 4);
        if (!(slo[k] == slo[k]) || !(shi[k] == shi[k]) || slo[k] - slo[k] != 0.f || shi[k] - shi[k] != 0.f)
            return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_mesh: non-finite vertex coordinate");
        ext = fmaxf(ext, shi[k] - slo[k]);
        amax = fmaxf(amax, fmaxf(fabsf(slo[k]), fabsf(shi[k])));
    }
    // Leaf boxes are padded by 2^-17 of the scene scale (0.19 mm for a 25 m room): float32 rounding of
    // the slab test (~1e-6 m) and of the Moller-Trumbore acceptance can then never cull a valid hit.
    const float pad = ldexpf(fmaxf(ext, amax), -17);
    // cell grid of the 32-byte node format: 65533 cells span the scene plus a margin of two pads on either side, so
    // every padded box quantises to [1, 65534] and the +-1 cell widening stays inside 16 bits
    NodeQ nq;
    for (int k = 0; k < 3; ++k) {
        const float span = (shi[k] - slo[k]) + 8.f * pad;
        nq.o[k] = slo[k] - 4.f * pad;
        nq.s[k] = fmaxf(span, 1e-6f) / 65533.f;
    }
    ctx->nodeq = nq;
    ctx->node_format = format;

    k_morton<<<gT, TB, 0, stream>>>(verts, tris, T, meta, k0, v0);
    LRC_CHECK_LAUNCH(ctx, "k_morton");
    uint64_t *kin = k0, *kout = k1;
    uint32_t *vin = v0, *vout = v1;
    for (int pass = 0; pass < 6; ++pass) {       // key bits 15..62 (k_morton clears the low 15)
        const int shift = 15 + 8 * pass;
        k_rs_hist<<<nb, RS_THREADS, 0, stream>>>(kin, T, shift, hist, nb);
        LRC_CHECK_LAUNCH(ctx, "k_rs_hist");
        k_rs_scan<<<1, 1024, 0, stream>>>(hist, (int64_t)256 * nb);
        LRC_CHECK_LAUNCH(ctx, "k_rs_scan");
        k_rs_scatter<<<nb, RS_THREADS, 0, stream>>>(kin, vin, kout, vout, T, shift, hist, nb);
        LRC_CHECK_LAUNCH(ctx, "k_rs_scatter");
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    // after an even number of passes the sorted data is back in (k0, v0) == (kin, vin)
    int ploc_iters = 0;
    if (quality) {
        // ---- SAH-quality build: PLOC over the Morton-ordered leaves (bvh_ploc.cuh) ----
        k_leaf_init<<<gT, TB, 0, stream>>>(verts, tris, T, vin, pad, nullptr, leaf_lo, leaf_hi);
        LRC_CHECK_LAUNCH(ctx, "k_leaf_init");
        if (!ctx->h_pin) LRC_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_pin, sizeof(int) * 16, cudaHostAllocDefault));
        int* cid[2] = {(int*)(base + o_cid0), (int*)(base + o_cid1)};
        float4* clo[2] = {(float4*)(base + o_clo0), (float4*)(base + o_clo1)};
        float4* chi[2] = {(float4*)(base + o_chi0), (float4*)(base + o_chi1)};
        PlocTree tr;
        tr.left = (int*)children; tr.right = tr.left + n_nodes;
        tr.cl = (int*)range; tr.start_node = tr.cl + n_nodes;
        tr.lo = node_lo; tr.hi = node_hi;
        tr.parent_node = parent_node; tr.parent_leaf = parent_leaf;
        tr.start_leaf = (int*)(base + o_sl);
        k_ploc_init<<<gT, TB, 0, stream>>>((int)T, leaf_lo, leaf_hi, cid[0], clo[0], chi[0]);
        LRC_CHECK_LAUNCH(ctx, "k_ploc_init");
        int R = (int)ctx->opt_ploc_radius;
        if ((rc = ploc_build(ctx, (int)T, R, cid, clo, chi, (int*)(base + o_nn), (unsigned char*)(base + o_role), (uint2*)(base + o_bsum),
                             tr, ctx->h_pin, &ploc_iters, stream))) return rc;
        LRC_CUDA(ctx, cudaMemsetAsync(parent_node + (T - 2), 0xff, sizeof(int), stream));      // the root (last node created) has no parent
        k_ploc_positions<<<(unsigned)((2 * T - 1 + TB - 1) / TB), TB, 0, stream>>>((int)T, tr, meta);
        LRC_CHECK_LAUNCH(ctx, "k_ploc_positions");
        k_ploc_emit<<<(unsigned)((T - 1 + TB - 1) / TB), TB, 0, stream>>>((int)T, tr, leaf_lo, leaf_hi, tmp_nodes, format, nq,
                                                                          (int)ctx->opt_leaf_size, meta, live);
        LRC_CHECK_LAUNCH(ctx, "k_ploc_emit");
        k_tri_records<<<gT, TB, 0, stream>>>(verts, tris, T, vin, tr.start_leaf, ctx->tris);
        LRC_CHECK_LAUNCH(ctx, "k_tri_records");
    } else {
        k_leaf_init<<<gT, TB, 0, stream>>>(verts, tris, T, vin, pad, ctx->tris, leaf_lo, leaf_hi);
        LRC_CHECK_LAUNCH(ctx, "k_leaf_init");
        if (T == 1) {
            k_single_leaf_root<<<1, 1, 0, stream>>>(leaf_lo, leaf_hi, ctx->nodes, meta, format, nq);
            LRC_CHECK_LAUNCH(ctx, "k_single_leaf_root");
        } else {
            const unsigned gN = (unsigned)((T - 1 + TB - 1) / TB);
            k_hierarchy<<<gN, TB, 0, stream>>>(kin, (int)T, children, parent_node, parent_leaf, range);
            LRC_CHECK_LAUNCH(ctx, "k_hierarchy");
            LRC_CUDA(ctx, cudaMemsetAsync(flags, 0, sizeof(int) * n_nodes, stream));
            k_refit<<<gT, TB, 0, stream>>>((int)T, parent_leaf, parent_node, children, leaf_lo, leaf_hi, node_lo, node_hi, flags,
                                           tmp_nodes, meta, format, nq, range, (int)ctx->opt_leaf_size);
            LRC_CHECK_LAUNCH(ctx, "k_refit");
            if (compact) {
                k_live_flags_lbvh<<<gN, TB, 0, stream>>>((int)n_nodes, range, (int)ctx->opt_leaf_size, live);
                LRC_CHECK_LAUNCH(ctx, "k_live_flags_lbvh");
            }
        }
    }
    int64_t n_live = n_nodes;
    if (T > 1 && compact) {
        // exclusive scan of the live flags = new record index; the total comes back with the build meta
        LRC_CUDA(ctx, cudaMemcpyAsync(new_index, live, sizeof(unsigned) * (size_t)n_nodes, cudaMemcpyDeviceToDevice, stream));
        k_rs_scan<<<1, 1024, 0, stream>>>(new_index, n_nodes);
        LRC_CHECK_LAUNCH(ctx, "k_rs_scan");
        k_compact_nodes<<<(unsigned)((n_nodes + TB - 1) / TB), TB, 0, stream>>>((int)n_nodes, tmp_nodes, new_index, live, ctx->nodes, format);
        LRC_CHECK_LAUNCH(ctx, "k_compact_nodes");
        unsigned tail[2];
        LRC_CUDA(ctx, cudaMemcpyAsync(&tail[0], new_index + (n_nodes - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
        LRC_CUDA(ctx, cudaMemcpyAsync(&tail[1], live + (n_nodes - 1), sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
        LRC_CUDA(ctx, cudaStreamSynchronize(stream));
        n_live = (int64_t)tail[0] + tail[1];
    }
    LRC_CUDA(ctx, cudaMemcpyAsync(&hm, meta, sizeof hm, cudaMemcpyDeviceToHost, stream));
    LRC_CUDA(ctx, cudaStreamSynchronize(stream));
    ctx->root = hm.root;
    ctx->build_quality = quality;
    ctx->ploc_iterations = ploc_iters;
    {
        const int K = (1 << LRC_TOP_LEVELS_MAX) - 1;
        if (!ctx->top_table) LRC_CUDA(ctx, cudaMalloc((void**)&ctx->top_table, sizeof(float4) * 4 * K));
        if (format == 0) {      // the shared-memory top-of-tree variant exists for the float format only
            k_top_table<<<1, 256, 0, stream>>>(ctx->nodes, ctx->top_table, K, ctx->root);
            LRC_CHECK_LAUNCH(ctx, "k_top_table");
        }
    }
    // scans that follow on other streams must see the finished tree; the scratch is theirs again after this point
    if (!ctx->scratch_event) LRC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->scratch_event, cudaEventDisableTiming));
    LRC_CUDA(ctx, cudaEventRecord(ctx->scratch_event, stream));

    lrc_bvh_info& inf = ctx->info;
    inf.num_tris = T;
    inf.num_nodes = n_live;
    inf.max_depth = hm.height;
    for (int k = 0; k < 3; ++k) { inf.scene_min[k] = slo[k]; inf.scene_max[k] = shi[k]; }
    inf.box_pad = pad;
    {
        float ex = hm.root_hi[0] - hm.root_lo[0], ey = hm.root_hi[1] - hm.root_lo[1], ez = hm.root_hi[2] - hm.root_lo[2];
        ctx->root_area = 2.0 * ((double)ex * ey + (double)ey * ez + (double)ez * ex);
        inf.sah_cost = -1.f;   // computed on demand by lrc_bvh_get_info
    }
    inf.bytes_nodes = (int64_t)sizeof(float4) * (format == 1 ? 2 : 4) * n_live;
    inf.bytes_tris = (int64_t)sizeof(float4) * 3 * T;
    if (hm.height + 1 >= LRC_STACK_DEPTH && quality) {
        // a PLOC tree has no height bound; the radix tree's is the key length -- rebuild with it rather than fail
        const int64_t keep = ctx->opt_build_quality;
        ctx->opt_build_quality = 0;
        rc = lrc_set_mesh(ctx, verts, V, tris, T, tri_label, stream_);
        ctx->opt_build_quality = keep;
        return rc;
    }
    if (hm.height + 1 >= LRC_STACK_DEPTH) {
        char buf[32];
        snprintf(buf, sizeof buf, "%d", hm.height);
        return lrc_fail(ctx, LRC_ERR_CAPACITY, "lrc_set_mesh: BVH height %s exceeds the traversal stack (degenerate mesh?)", buf);
    }
    ctx->has_mesh = true;
    return LRC_OK;
}

extern "C" int lrc_bvh_get_info(lrc_ctx* ctx, lrc_bvh_info* h_info)
{
    if (!ctx || !h_info) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_bvh_get_info: NULL argument");
    if (!ctx->has_mesh) return lrc_fail(ctx, LRC_ERR_NO_MESH, "lrc_bvh_get_info: no mesh set");
    if (ctx->info.sah_cost < 0.f && ctx->T > 0) {
        // SAH cost = (area(root) + sum of all child-box areas) / area(root), C_traverse = C_intersect = 1
        LRC_CUDA(ctx, cudaSetDevice(ctx->device));
        double* d_sum = nullptr;
        LRC_CUDA(ctx, cudaMalloc((void**)&d_sum, sizeof(double)));
        LRC_CUDA(ctx, cudaMemset(d_sum, 0, sizeof(double)));
        k_sah_sum<<<296, 256>>>(ctx->nodes, ctx->info.num_nodes, d_sum, ctx->node_format, ctx->nodeq);
        LRC_CHECK_LAUNCH(ctx, "k_sah_sum");
        double h_sum = 0.0;
        LRC_CUDA(ctx, cudaMemcpy(&h_sum, d_sum, sizeof(double), cudaMemcpyDeviceToHost));
        LRC_CUDA(ctx, cudaFree(d_sum));
        ctx->info.sah_cost = ctx->root_area > 0 ? (float)((ctx->root_area + h_sum) / ctx->root_area) : 0.f;
    }
    *h_info = ctx->info;
    return LRC_OK;
}
