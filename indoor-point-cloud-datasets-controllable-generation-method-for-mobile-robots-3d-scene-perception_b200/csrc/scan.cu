// scan.cu -- ray generation, traversal launches (with the frame arithmetic: point reconstruction, range filter,
// incident angle) and the ordered compaction with label gather.  Together these replace, per waypoint,
//     RaycastEngineCPU.lidar_intersect_mesh / rays_intersect_mesh  (reference raycast_engine_cpu.py:24-111)
// and IndoorLidar.get_rays / DualAxisLidar.get_rays                 (reference indoor_lidar.py:27-131,224-319).
#include "traverse.cuh"

char g_lrc_global_err[512] = {0};

namespace {

constexpr double PI_D = 3.141592653589793;

enum { MODE_SINGLE = 0, MODE_DUAL = 1, MODE_RAYS = 2 };

struct RayGen {
    // common
    const double* poses;       // P x 16, row-major 4x4
    int64_t pose0;             // first pose of this launch (index into poses)
    int N;                     // rays per frame (dense, before dropout)
    int W;                     // single: horizontal_res ; dual: points_per_line
    int H;                     // single: lines ; dual: num_lines
    // single axis: tables [cos beta | sin beta | cos alpha | sin alpha] (float64)
    const double* tab;
    int uniform_mode;          // 1: round the local direction to float32 before rotating (indoor_lidar.py:82)
    // dual axis
    double theta_min, theta_max, tstep, pstep, swing_amp, swing_freq;
    // noise
    double angle_std, dropout_p, range_std;
    uint32_t k0, k1;
    uint64_t pose_index_base;
    // explicit rays
    const float* rays;
};

struct Ray { float ox, oy, oz, dx, dy, dz; bool keep; };

__device__ __forceinline__ void rotate(const double* __restrict__ M, double lx, double ly, double lz, Ray& r)
{
    // world = R . local ; origin = pose[:3,3] ; float64 math, a single rounding to float32
    r.dx = (float)(M[0] * lx + M[1] * ly + M[2] * lz);
    r.dy = (float)(M[4] * lx + M[5] * ly + M[6] * lz);
    r.dz = (float)(M[8] * lx + M[9] * ly + M[10] * lz);
    r.ox = (float)M[3];
    r.oy = (float)M[7];
    r.oz = (float)M[11];
}

// ray index inside a frame = j*W + i (line-major, azimuth-minor): reference indoor_lidar.py:108-110
__device__ __forceinline__ Ray gen_single(const RayGen& g, int64_t pose, int r)
{
    const int j = r / g.W, i = r - j * g.W;
    const double cb = g.tab[i], sb = g.tab[g.W + i], ca = g.tab[2 * g.W + j], sa = g.tab[2 * g.W + g.H + j];
    double lx = ca * cb, ly = ca * sb, lz = sa;
    if (g.uniform_mode) { lx = (double)(float)lx; ly = (double)(float)ly; lz = (double)(float)lz; }
    Ray out;
    rotate(g.poses + 16 * (g.pose0 + pose), lx, ly, lz, out);
    out.keep = true;
    return out;
}

// reference indoor_lidar.py:247-294
__device__ __forceinline__ Ray gen_dual(const RayGen& g, int64_t pose, int r)
{
    const int line = r / g.W, k = r - line * g.W;
    const double base = (g.H > 1 && line == g.H - 1) ? g.theta_min : g.theta_max + (double)line * g.tstep;
    const double phase = (double)line * PI_D / (double)g.H;
    double phi = (double)k * g.pstep;
    double theta = base + g.swing_amp * sin(g.swing_freq * phi + phase);
    theta = fmin(fmax(theta, g.theta_min), g.theta_max);
    Ray out;
    out.keep = true;
    if (g.angle_std > 0.0 || g.dropout_p > 0.0) {
        const uint64_t pidx = g.pose_index_base + (uint64_t)(g.pose0 + pose);
        const uint4 rnd = philox4x32_10(make_uint4((uint32_t)r, (uint32_t)pidx, (uint32_t)(pidx >> 32), 0u), g.k0, g.k1);
        if (g.angle_std > 0.0) {   // Box-Muller; noise is added AFTER the clip (indoor_lidar.py:267-272)
            const double rad = sqrt(-2.0 * log(u01(rnd.x)));
            double sn, cs;
            sincos(2.0 * PI_D * u01(rnd.y), &sn, &cs);
            phi += g.angle_std * (rad * cs);
            theta += g.angle_std * (rad * sn);
        }
        if (g.dropout_p > 0.0) out.keep = u01(rnd.z) > g.dropout_p;   // indoor_lidar.py:292-294
    }
    double st, ct, sp, cp;
    sincos(theta, &st, &ct);
    sincos(phi, &sp, &cp);
    rotate(g.poses + 16 * (g.pose0 + pose), ct * cp, ct * sp, st, out);
    return out;
}

__device__ __forceinline__ Ray gen_explicit(const RayGen& g, int64_t idx)
{
    const float2* p = reinterpret_cast<const float2*>(g.rays + 6 * idx);
    const float2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    Ray out;
    out.ox = a.x; out.oy = a.y; out.oz = b.x;
    out.dx = b.y; out.dy = c.x; out.dz = c.y;
    out.keep = true;
    return out;
}

template <int MODE>
__device__ __forceinline__ Ray gen_ray(const RayGen& g, int64_t idx, int64_t& pose, int& r)
{
    if (MODE == MODE_RAYS) { pose = 0; r = (int)idx; return gen_explicit(g, idx); }
    pose = idx / g.N;
    r = (int)(idx - pose * g.N);
    return MODE == MODE_SINGLE ? gen_single(g, pose, r) : gen_dual(g, pose, r);
}

// standard normal for the range noise: Philox stream 1 of the same (seed, pose, ray) counter
__device__ __forceinline__ double range_normal(const RayGen& g, int64_t pose, int r)
{
    const uint64_t pidx = g.pose_index_base + (uint64_t)(g.pose0 + pose);
    const uint4 rnd = philox4x32_10(make_uint4((uint32_t)r, (uint32_t)pidx, (uint32_t)(pidx >> 32), 1u), g.k0, g.k1);
    return sqrt(-2.0 * log(u01(rnd.x))) * cos(2.0 * PI_D * u01(rnd.y));
}

// ---- table kernel: cos/sin of azimuth and elevation in float64 ---------------------------------------
// beta = -(i - W/2)/W*2*pi (indoor_lidar.py:113), alpha = deg2rad(v[j]) (:116); uniform mode: linspace tables (:66-72)
__global__ void k_single_tables(double* tab, int W, int H, const double* vdeg, int uniform_mode, double fov_up_deg,
                                double fov_down_deg)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < W) {
        double beta;
        if (uniform_mode) beta = (double)t * ((2.0 * PI_D - 0.0) / (double)W);
        else beta = -((double)t - (double)W / 2.0) / (double)W * 2.0 * PI_D;
        tab[t] = cos(beta);
        tab[W + t] = sin(beta);
    } else if (t < W + H) {
        int j = t - W;
        double alpha;
        if (uniform_mode) {
            double up = fov_up_deg * (PI_D / 180.0), dn = fov_down_deg * (PI_D / 180.0);
            double step = H > 1 ? ((-dn) - up) / (double)(H - 1) : 0.0;
            alpha = (H > 1 && j == H - 1) ? -dn : up + (double)j * step;
        } else alpha = vdeg[j] * (PI_D / 180.0);
        tab[2 * W + j] = cos(alpha);
        tab[2 * W + H + j] = sin(alpha);
    }
}

// ---- ray table export (get_rays) ---------------------------------------------------------------------
template <int MODE>
__global__ void k_gen_rays(RayGen g, int64_t n, float* __restrict__ rays, uint8_t* __restrict__ keep)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    int64_t pose; int r;
    Ray ray = gen_ray<MODE>(g, idx, pose, r);
    float2* o = reinterpret_cast<float2*>(rays + 6 * idx);
    o[0] = make_float2(ray.ox, ray.oy);
    o[1] = make_float2(ray.oz, ray.dx);
    o[2] = make_float2(ray.dy, ray.dz);
    if (keep) keep[idx] = ray.keep ? 1 : 0;
}

// ---- traversal kernels -------------------------------------------------------------------------------
// Frame-level arithmetic that rides in the traversal kernel (the float64 pipe is otherwise idle there):
//     d^ = d / sqrt((dx*dx + dy*dy) + dz*dz) ; p = o + d^ * t          float32, separate roundings (raycast_engine_cpu.py:57,62)
//     dist = sqrt(((p - c)^2).sum()) in float64 from the float32 point ; keep <=> dist < max_range   (strict, :95-97)
//     incident = degrees(arccos(|(p - c)_z / dist|))                   float64 (:100-107)
struct FrameMath {
    double max_range;          // < 0: no range filter, no incident angle (rays_intersect_mesh)
    double cx, cy, cz;         // frame centre for MODE_RAYS; scan modes read pose[:3,3]
    int wire;                  // compact wire format: the 8-byte scratch slot of the angle carries the float32 hit distance t instead
};

// One ray's share of the frame arithmetic (see above): writes the scratch record of ray idx, returns whether it is kept.
template <int MODE, bool CS = false>
__device__ __forceinline__ bool frame_epilogue(const RayGen& g, const FrameMath& fm, int64_t idx, int64_t pose, int r, float ox, float oy,
                                               float oz, float dx, float dy, float dz, float t, uint32_t id, float4* __restrict__ hp,
                                               double* __restrict__ inc_out)
{
    bool keep = false;
    float4 o = make_float4(0.f, 0.f, 0.f, __uint_as_float(LRC_MISS_ID));
    double inc = 0.0;
    if (id != LRC_MISS_ID) {
        if (g.range_std > 0.0) t = __fadd_rn(t, (float)(g.range_std * range_normal(g, pose, r)));
        float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        o.x = __fadd_rn(ox, __fmul_rn(__fdiv_rn(dx, nrm), t));
        o.y = __fadd_rn(oy, __fmul_rn(__fdiv_rn(dy, nrm), t));
        o.z = __fadd_rn(oz, __fmul_rn(__fdiv_rn(dz, nrm), t));
        keep = true;
        if (fm.max_range >= 0.0) {
            double cx = fm.cx, cy = fm.cy, cz = fm.cz;
            if (MODE != MODE_RAYS) {
                const double* M = g.poses + 16 * (g.pose0 + pose);
                cx = M[3]; cy = M[7]; cz = M[11];
            }
            const double ddx = __dsub_rn((double)o.x, cx), ddy = __dsub_rn((double)o.y, cy), ddz = __dsub_rn((double)o.z, cz);
            const double dist = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)), __dmul_rn(ddz, ddz)));
            keep = dist < fm.max_range;
            if (inc_out && !fm.wire) inc = __dmul_rn(acos(fabs(__ddiv_rn(ddz, dist))), 180.0 / PI_D);
        }
        if (keep) o.w = __uint_as_float(id);
    }
    if (fm.wire) inc = __longlong_as_double((long long)__float_as_uint(t));      // t after range noise: what p was built from
    if (CS) {      // evict-first: the scratch is read once by k_compact and should not push BVH lines out of L2
        __stcs(hp + idx, o);
        if (inc_out) __stcs(inc_out + idx, inc);
    } else {
        hp[idx] = o;
        if (inc_out) inc_out[idx] = inc;
    }
    return keep;
}

constexpr int TRACE_THREADS = 128;

// OUT_DENSE: write (t_hit, prim_id) per ray (cast_rays).  Otherwise, per ray, the float32 hit point + triangle id
// (float4; id = MISS when the ray missed, was dropped or failed the range test), the incident angle, and per BLOCK
// the number of kept rays (block_count) -- the first stage of the ordered compaction.
// One ray of a k_trace launch: generate, traverse, frame arithmetic (or the dense t / id outputs).  Returns "kept".
template <int MODE, bool COUNT, bool OUT_DENSE, int VARIANT>
__device__ __forceinline__ bool trace_one(const RayGen& g, const FrameMath& fm, const float4* __restrict__ nodes,
                                          const float4* __restrict__ tris, int64_t idx, int has_tris, float4* __restrict__ hp,
                                          double* __restrict__ inc_out, float* __restrict__ t_hit, uint32_t* __restrict__ prim_id,
                                          const float4* s_top, int top_n, int stack_levels, const NodeQ& nq, int root, unsigned& nn,
                                          unsigned& nt, unsigned& nr, unsigned& nh, cudaTextureObject_t ntex = 0)
{
    int64_t pose; int r;
    Ray ray = gen_ray<MODE>(g, idx, pose, r);
    float t = LRC_INF;
    uint32_t id = LRC_MISS_ID;
    if (ray.keep && has_tris) {
        trace_ray<VARIANT, COUNT>(nodes, tris, s_top, top_n, stack_levels, nq, root, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, t, id, nn, nt, ntex);
        nr += 1;
        nh += id != LRC_MISS_ID;
    }
    if (OUT_DENSE) {
        t_hit[idx] = t;
        prim_id[idx] = id;
        return false;
    }
    return frame_epilogue<MODE, (VARIANT & 4096) != 0>(g, fm, idx, pose, r, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, t, id, hp, inc_out);
}

// VARIANT bit 8 ("persistent"): the grid is sized to fill the machine once and every WARP keeps fetching work -- one
// 128-ray block of the compaction at a time (an atomic ticket per block), walked as tiles of 32 consecutive rays.  No
// block-wide barrier parks finished warps behind the block's slowest ray (ncu: ~18 % of the resident warp-time is spent at
// that barrier) and the tail of the launch is balanced at warp granularity.  The warp owns its whole block, so the keep
// count is a plain store.
template <int MODE, bool COUNT, bool OUT_DENSE, int VARIANT>
__global__ void __launch_bounds__(TRACE_THREADS, (VARIANT & 4) ? 16 : (VARIANT & 512) ? 10 : 12)   // 40 registers (48 warps per SM); bit 2: 32 registers (64 warps); bit 9: 48 registers (40 warps)
k_trace(RayGen g, FrameMath fm, const float4* __restrict__ nodes, const float4* __restrict__ tris, int64_t n, int has_tris,
        float4* __restrict__ hp, double* __restrict__ inc_out, unsigned* __restrict__ block_count, float* __restrict__ t_hit,
        uint32_t* __restrict__ prim_id, unsigned long long* counters, const float4* __restrict__ top_table, int top_n,
        int stack_levels, const __grid_constant__ NodeQ nq, int root, int tile_shift, cudaTextureObject_t ntex)
{
    extern __shared__ float4 s_top[];
    constexpr bool PERSIST = (VARIANT & 256) != 0;
    constexpr bool NOBAR = (VARIANT & 2048) != 0 && !PERSIST && !OUT_DENSE;
    __shared__ unsigned s_keep, s_done;
    if (NOBAR) {          // all warps of the block are still together here: this barrier costs nothing
        if (threadIdx.x == 0) { s_keep = 0u; s_done = 0u; }
        __syncthreads();
    }
    if (VARIANT & 8) {      // stage the top of the tree (heap order, built by lrc_set_mesh) in shared memory
        for (int i = threadIdx.x; i < 4 * top_n; i += blockDim.x) s_top[i] = __ldg(top_table + i);
        __syncthreads();
    }
    unsigned nn = 0, nt = 0, nr = 0, nh = 0;
    if (PERSIST) {
        unsigned* ticket = reinterpret_cast<unsigned*>(counters + 4);
        const unsigned tiles_per_block = 1u << tile_shift;
        const unsigned n_blocks = (unsigned)((n + (32u << tile_shift) - 1) >> (5 + tile_shift));
        for (;;) {
            unsigned b = 0;
            if (lane_id() == 0) b = atomicAdd(ticket, 1u);
            b = __shfl_sync(0xffffffffu, b, 0);
            if (b >= n_blocks) break;
            unsigned kept = 0;
            for (unsigned k = 0; k < tiles_per_block; ++k) {
                const int64_t idx = (((int64_t)b << tile_shift) + k) * 32 + lane_id();
                bool keep = false;
                if (idx < n) keep = trace_one<MODE, COUNT, OUT_DENSE, VARIANT>(g, fm, nodes, tris, idx, has_tris, hp, inc_out, t_hit, prim_id, s_top, top_n, stack_levels, nq, root, nn, nt, nr, nh, ntex);
                kept += __popc(__ballot_sync(0xffffffffu, keep));
            }
            if (!OUT_DENSE && lane_id() == 0) block_count[b] = kept;
        }
    } else {
        const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        bool keep = false;
        if (idx < n) keep = trace_one<MODE, COUNT, OUT_DENSE, VARIANT>(g, fm, nodes, tris, idx, has_tris, hp, inc_out, t_hit, prim_id, s_top, top_n, stack_levels, nq, root, nn, nt, nr, nh, ntex);
        if (NOBAR) {
            // per-warp keep counts meet in shared memory; the warp that arrives last publishes the block's count.  No warp
            // waits for the block's slowest ray: its slot is free for the next block's warps as soon as it is done.
            const unsigned c = __popc(__ballot_sync(0xffffffffu, keep));
            if (lane_id() == 0) {
                atomicAdd(&s_keep, c);
                __threadfence_block();
                if (atomicAdd(&s_done, 1u) == (blockDim.x >> 5) - 1u) {
                    __threadfence_block();
                    block_count[blockIdx.x] = atomicAdd(&s_keep, 0u);
                }
            }
        } else if (!OUT_DENSE) {
            const int c = __syncthreads_count(keep ? 1 : 0);
            if (threadIdx.x == 0) block_count[blockIdx.x] = (unsigned)c;
        }
    }
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nn += __shfl_xor_sync(0xffffffffu, nn, o);
            nt += __shfl_xor_sync(0xffffffffu, nt, o);
            nr += __shfl_xor_sync(0xffffffffu, nr, o);
            nh += __shfl_xor_sync(0xffffffffu, nh, o);
        }
        if (lane_id() == 0) {
            atomicAdd(&counters[0], (unsigned long long)nr);
            atomicAdd(&counters[1], (unsigned long long)nn);
            atomicAdd(&counters[2], (unsigned long long)nt);
            atomicAdd(&counters[3], (unsigned long long)nh);
        }
    }
}

// K adjacent rays per thread (traverse.cuh "thread packets"): block = 128 threads = 128 K rays = K blocks of the compaction.
template <int MODE, bool COUNT, int K>
__global__ void __launch_bounds__(TRACE_THREADS, K == 2 ? 8 : 5)
k_trace_k(RayGen g, FrameMath fm, const float4* __restrict__ nodes, const float4* __restrict__ tris, int64_t n, int has_tris,
          float4* __restrict__ hp, double* __restrict__ inc_out, unsigned* __restrict__ block_count, unsigned long long* counters, int root)
{
    __shared__ unsigned s_cnt[K];
    if (threadIdx.x < K) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t base = ((int64_t)blockIdx.x * TRACE_THREADS + threadIdx.x) * K;
    Packet<K> pk;
    unsigned nn = 0, nt = 0, nr = 0, nh = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int64_t idx = base + k;
        if (idx < n) {
            int64_t pose; int r;
            const Ray ray = gen_ray<MODE>(g, idx, pose, r);
            const bool live = ray.keep && has_tris;
            packet_set_ray<K>(pk, k, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, live);
            nr += live;
        } else {
            packet_set_ray<K>(pk, k, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, false);
        }
    }
    if (base < n && has_tris) trace_packet<K, COUNT>(nodes, tris, root, pk, nn, nt);
    unsigned kept = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int64_t idx = base + k;
        if (idx < n) {
            int64_t pose = 0; int r = (int)idx;
            if (MODE != MODE_RAYS) { pose = idx / g.N; r = (int)(idx - pose * g.N); }
            const uint32_t id = pk.best_id[k];
            nh += id != LRC_MISS_ID;
            kept += frame_epilogue<MODE>(g, fm, idx, pose, r, pk.ox[k], pk.oy[k], pk.oz[k], pk.dx[k], pk.dy[k], pk.dz[k], pk.best_t[k], id, hp, inc_out);
        }
    }
    // keep counts per 128 rays: sub-block j of this block = threads [j * 128 / K, (j + 1) * 128 / K), i.e. whole warps
    const unsigned wsum = __reduce_add_sync(0xffffffffu, kept);
    if (lane_id() == 0 && wsum) atomicAdd(&s_cnt[(threadIdx.x * K) / TRACE_THREADS], wsum);
    __syncthreads();
    if (threadIdx.x < K && ((int64_t)blockIdx.x * K + threadIdx.x) * TRACE_THREADS < n) block_count[(int64_t)blockIdx.x * K + threadIdx.x] = s_cnt[threadIdx.x];
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nn += __shfl_xor_sync(0xffffffffu, nn, o);
            nt += __shfl_xor_sync(0xffffffffu, nt, o);
            nr += __shfl_xor_sync(0xffffffffu, nr, o);
            nh += __shfl_xor_sync(0xffffffffu, nh, o);
        }
        if (lane_id() == 0) {
            atomicAdd(&counters[0], (unsigned long long)nr);
            atomicAdd(&counters[1], (unsigned long long)nn);
            atomicAdd(&counters[2], (unsigned long long)nt);
            atomicAdd(&counters[3], (unsigned long long)nh);
        }
    }
}

// Warp packets (traverse.cuh trace_warp): one ray per thread as in k_trace, but the warp walks the tree together with
// one shared-memory stack per warp.  Scan modes only -- explicit rays carry no promise of coherence.
template <int MODE, bool COUNT>
__global__ void __launch_bounds__(TRACE_THREADS, 12)
k_trace_w(RayGen g, FrameMath fm, const float4* __restrict__ nodes, const float4* __restrict__ tris, int64_t n, int has_tris,
          float4* __restrict__ hp, double* __restrict__ inc_out, unsigned* __restrict__ block_count, unsigned long long* counters, int root)
{
    __shared__ int s_link[TRACE_THREADS / 32][LRC_WSTACK];
    __shared__ unsigned s_t[TRACE_THREADS / 32][LRC_WSTACK];
    const int w = threadIdx.x >> 5;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned nn = 0, nt = 0, nr = 0, nh = 0;
    int64_t pose = 0; int r = 0;
    Ray ray;
    ray.ox = ray.oy = ray.oz = 0.f; ray.dx = 1.f; ray.dy = ray.dz = 0.f; ray.keep = false;
    if (idx < n) ray = gen_ray<MODE>(g, idx, pose, r);
    const bool live = idx < n && ray.keep && has_tris;
    float t; uint32_t id;
    trace_warp<COUNT>(nodes, tris, root, s_link[w], s_t[w], ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, live, t, id, nn, nt);
    if (!live) { t = LRC_INF; id = LRC_MISS_ID; }
    nr += live;
    nh += id != LRC_MISS_ID;
    bool keep = false;
    if (idx < n) keep = frame_epilogue<MODE>(g, fm, idx, pose, r, ray.ox, ray.oy, ray.oz, ray.dx, ray.dy, ray.dz, t, id, hp, inc_out);
    const int c = __syncthreads_count(keep ? 1 : 0);
    if (threadIdx.x == 0) block_count[blockIdx.x] = (unsigned)c;
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nn += __shfl_xor_sync(0xffffffffu, nn, o);
            nt += __shfl_xor_sync(0xffffffffu, nt, o);
            nr += __shfl_xor_sync(0xffffffffu, nr, o);
            nh += __shfl_xor_sync(0xffffffffu, nh, o);
        }
        if (lane_id() == 0) {
            atomicAdd(&counters[0], (unsigned long long)nr);
            atomicAdd(&counters[1], (unsigned long long)nn);
            atomicAdd(&counters[2], (unsigned long long)nt);
            atomicAdd(&counters[3], (unsigned long long)nh);
        }
    }
}

// exhaustive validation kernel: every ray against every triangle record
__global__ void k_brute(const float* __restrict__ rays, int64_t n, const float4* __restrict__ tris, int64_t T,
                        float* __restrict__ t_hit, uint32_t* __restrict__ prim_id)
{
    int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const float* p = rays + 6 * idx;
    float ox = p[0], oy = p[1], oz = p[2], dx = p[3], dy = p[4], dz = p[5];
    float best_t = LRC_INF;
    uint32_t best_id = LRC_MISS_ID;
    for (int64_t s = 0; s < T; ++s) {
        float4 v0 = __ldg(tris + 3 * s), e1 = __ldg(tris + 3 * s + 1), e2 = __ldg(tris + 3 * s + 2);
        float t;
        if (mt_hit(ox, oy, oz, dx, dy, dz, v0, e1, e2, t)) {
            uint32_t id = __float_as_uint(v0.w);
            if (t < best_t || (t == best_t && id < best_id)) { best_t = t; best_id = id; }
        }
    }
    t_hit[idx] = best_t;
    prim_id[idx] = best_id;
}

// ---- ordered compaction: stage 2 (scan of block counts) and stage 3 (streaming scatter) ------------------
// The reference's boolean-mask indexing keeps ray order (raycast_engine_cpu.py:71,:97), so the compaction must be
// order-preserving: per-block keep counts (written by k_trace) -> exclusive scan -> each block scatters its kept
// rays at base + warp-ballot prefix.  No atomics decide positions, so the output order is deterministic.
constexpr int SC_ITEMS = 16;   // counts per thread per trip (4 x uint4): 16384 counts per trip of the single block
__global__ void __launch_bounds__(1024) k_scan_counts(const unsigned* __restrict__ counts, unsigned* __restrict__ base,
                                                      int64_t nb, const long long* __restrict__ run_in,
                                                      long long* __restrict__ run_out, int64_t* __restrict__ frame_offset_last)
{
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (int64_t b0 = 0; b0 < nb; b0 += 1024 * SC_ITEMS) {
        const int64_t i0 = b0 + (int64_t)threadIdx.x * SC_ITEMS;
        unsigned v[SC_ITEMS];
        if (i0 + SC_ITEMS <= nb) {      // the arrays are 256 B aligned and i0 is a multiple of 16
#pragma unroll
            for (int k = 0; k < SC_ITEMS / 4; ++k) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(counts + i0) + k);
                v[4 * k] = q.x; v[4 * k + 1] = q.y; v[4 * k + 2] = q.z; v[4 * k + 3] = q.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < SC_ITEMS; ++k) v[k] = i0 + k < nb ? counts[i0 + k] : 0u;
        }
        unsigned mine = 0u;
#pragma unroll
        for (int k = 0; k < SC_ITEMS; ++k) mine += v[k];
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
        if (i0 + SC_ITEMS <= nb) {
#pragma unroll
            for (int k = 0; k < SC_ITEMS / 4; ++k) {
                uint4 q;
                q.x = run; run += v[4 * k];
                q.y = run; run += v[4 * k + 1];
                q.z = run; run += v[4 * k + 2];
                q.w = run; run += v[4 * k + 3];
                reinterpret_cast<uint4*>(base + i0)[k] = q;
            }
        } else {
#pragma unroll
            for (int k = 0; k < SC_ITEMS; ++k) {
                if (i0 + k < nb) base[i0 + k] = run;
                run += v[k];
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) carry = run;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const long long total = *run_in + (long long)carry;
        *run_out = total;
        if (frame_offset_last) *frame_offset_last = total;
    }
}

struct CompactParams {
    const float4* hp;
    const double* inc;
    const unsigned* base;      // exclusive scan of the per-block keep counts of this launch
    const long long* run_in;   // points emitted by earlier launches of this call
    int64_t n;                 // rays in this launch
    int64_t ray0;              // global index of the first ray of this launch
    int64_t N;                 // rays per frame
    int64_t total_rays;        // rays of the whole call (the owner of the last one writes the closing offset)
    int64_t P;                 // frames of the whole call
    const uint32_t* labels;
    lrc_out out;
    float* wire_t;             // compact wire format: hit distance and ray index of every kept point (this rank's region of its
    uint32_t* wire_ray;        // own gather buffer, indexed like out.xyz); the scratch slot `inc` carries t then
};

// Launched with the block size k_trace used (one keep count per block of rays).
__global__ void __launch_bounds__(TRACE_THREADS) k_compact(CompactParams q)
{
    __shared__ int s_warp[TRACE_THREADS / 32];
    const long long run0 = *q.run_in;
    const unsigned blk_base = q.base[blockIdx.x];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float4 h = make_float4(0.f, 0.f, 0.f, __uint_as_float(LRC_MISS_ID));
    double inc = 0.0;
    if (i < q.n) {
        h = __ldcs(q.hp + i);
        if (q.inc) inc = __ldcs(q.inc + i);
    }
    const uint32_t id = __float_as_uint(h.w);
    const bool keep = id != LRC_MISS_ID;
    const unsigned ballot = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) s_warp[w] = __popc(ballot);
    __syncthreads();
    int warp_off = 0;
#pragma unroll
    for (int k = 0; k < TRACE_THREADS / 32; ++k)
        if (k < w) warp_off += s_warp[k];
    const long long pos = run0 + blk_base + warp_off + __popc(ballot & ((1u << lane) - 1u));
    if (i < q.n) {
        const int64_t gidx = q.ray0 + i;
        const int64_t frame = gidx / q.N;
        const int64_t r = gidx - frame * q.N;
        if (r == 0) q.out.frame_offset[frame] = pos;
        if (keep && pos < q.out.capacity) {
            q.out.xyz[3 * pos + 0] = h.x;
            q.out.xyz[3 * pos + 1] = h.y;
            q.out.xyz[3 * pos + 2] = h.z;
            if (q.wire_t) {
                q.wire_t[pos] = __uint_as_float((unsigned)__double_as_longlong(inc));
                q.wire_ray[pos] = (uint32_t)r;
            } else if (q.out.incident_deg) q.out.incident_deg[pos] = inc;
            if (q.out.prim_id) q.out.prim_id[pos] = id;
            if (q.out.label) q.out.label[pos] = q.labels ? __ldg(q.labels + id) : 0u;
            if (q.out.ray_idx) q.out.ray_idx[pos] = (uint32_t)r;
        }
    }
}

#include "exchange.cuh"   // k_push, k_push_tma, k_wire_* (multi-GPU exchange kernels)

// ---- host-side launch plumbing -----------------------------------------------------------------------
int ensure_counters(lrc_ctx* ctx)
{
    if (!ctx->d_counters) {
        LRC_CUDA(ctx, cudaMalloc((void**)&ctx->d_counters, 8 * sizeof(unsigned long long)));      // [0..3] work counters, [4] persistent-kernel ticket
        LRC_CUDA(ctx, cudaMemset(ctx->d_counters, 0, 8 * sizeof(unsigned long long)));
    }
    return LRC_OK;
}

// rays per thread of the scan kernels: > 1 needs the paired node records (format 2) and 128-ray compaction blocks
inline int packet_rays(const lrc_ctx* ctx) { return ctx->node_format == 2 ? (int)ctx->opt_rays_per_thread : 1; }
inline bool warp_packets(const lrc_ctx* ctx) { return ctx->opt_warp_packet && ctx->node_format == 2 && packet_rays(ctx) == 1; }
// Threads per traversal block (= rays per keep count of the compaction).  "block" = 0 (default) picks by the size of the
// call: a one-frame call of 16 000 rays is 125 blocks of 128 threads on 148 SMs -- one warp per scheduler and nothing to hide
// latency with -- so small calls use smaller blocks (measured: tools/small_frame_block.py); trajectories use 128.
inline int scan_block_threads(const lrc_ctx* ctx, int64_t total_rays)
{
    if (packet_rays(ctx) > 1 || warp_packets(ctx)) return TRACE_THREADS;
    if (ctx->opt_block) return (int)ctx->opt_block;
    return total_rays <= (int64_t)1 << 16 ? 32 : total_rays <= (int64_t)1 << 18 ? 64 : TRACE_THREADS;
}

template <int MODE, bool DENSE>
int launch_trace(lrc_ctx* ctx, const RayGen& g, const FrameMath& fm, int64_t n, float4* hp, double* inc, unsigned* block_count,
                 float* t_hit, uint32_t* prim, cudaStream_t stream)
{
    if (n <= 0) return LRC_OK;
    const int has_tris = ctx->T > 0;
    if (!DENSE && packet_rays(ctx) > 1) {
        // K adjacent rays per thread (paired node records only)
        const int K = packet_rays(ctx);
        const unsigned pgrid = (unsigned)((n + (int64_t)TRACE_THREADS * K - 1) / ((int64_t)TRACE_THREADS * K));
#define LRC_LAUNCH_PACKET(COUNT, KK) \
        k_trace_k<MODE, COUNT, KK><<<pgrid, TRACE_THREADS, 0, stream>>>(g, fm, ctx->nodes, ctx->tris, n, has_tris, hp, inc, block_count, ctx->d_counters, ctx->root)
        if (ctx->counting) { if (K == 2) LRC_LAUNCH_PACKET(true, 2); else LRC_LAUNCH_PACKET(true, 4); }
        else { if (K == 2) LRC_LAUNCH_PACKET(false, 2); else LRC_LAUNCH_PACKET(false, 4); }
#undef LRC_LAUNCH_PACKET
        LRC_CHECK_LAUNCH(ctx, "k_trace_k");
        return LRC_OK;
    }
    if (!DENSE && MODE != MODE_RAYS && warp_packets(ctx)) {
        const unsigned wgrid = (unsigned)((n + TRACE_THREADS - 1) / TRACE_THREADS);
        if (ctx->counting) k_trace_w<MODE, true><<<wgrid, TRACE_THREADS, 0, stream>>>(g, fm, ctx->nodes, ctx->tris, n, has_tris, hp, inc, block_count, ctx->d_counters, ctx->root);
        else k_trace_w<MODE, false><<<wgrid, TRACE_THREADS, 0, stream>>>(g, fm, ctx->nodes, ctx->tris, n, has_tris, hp, inc, block_count, ctx->d_counters, ctx->root);
        LRC_CHECK_LAUNCH(ctx, "k_trace_w");
        return LRC_OK;
    }
    const int TB = DENSE ? TRACE_THREADS : ctx->cur_block;
    unsigned grid = (unsigned)((n + TB - 1) / TB);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cudaLaunchAttribute attr[1];
    cfg.attrs = attr;
    if (ctx->opt_l2_persist && ctx->l2_persist_max > 0 && ctx->bvh_bytes > 0) {
        // persisting lines for the BVH records; the window may be smaller than the tree (then its head -- the nodes --
        // is covered first) and the set-aside smaller than the window (then hitRatio picks that fraction of lines)
        const size_t win = ctx->bvh_bytes < ctx->l2_window_max ? ctx->bvh_bytes : ctx->l2_window_max;
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = ctx->bvh_block;
        attr[0].val.accessPolicyWindow.num_bytes = win;
        attr[0].val.accessPolicyWindow.hitRatio = win <= ctx->l2_persist_max ? 1.0f : (float)((double)ctx->l2_persist_max / (double)win);
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
        cfg.numAttrs = 1;
    }
    const float4* nodes = ctx->nodes;
    const float4* tris = ctx->tris;
    unsigned long long* counters = ctx->d_counters;
    const float4* top_table = ctx->top_table;
    const NodeQ nq = ctx->nodeq;
    const int root = ctx->root;
    const cudaTextureObject_t ntex = ctx->nodes_tex;
    // the resident tree's record format selects the kernel family; the variant option tunes the float-format kernels
    int64_t variant = ctx->node_format == 1 ? 37 : ctx->node_format == 2 ? ((ctx->opt_variant & 64) ? 193 : 129) : ctx->opt_variant;
    if (variant == 193 && ctx->opt_tune) {
        variant |= (ctx->opt_tune & 7) << 10;     // bits 10..12: prefetch, no barrier, streaming stores
        if (ctx->nodes_tex) variant |= ((ctx->opt_tune >> 3) & 3) << 13;      // bits 13..14: record halves fetched through the texture pipe
    }
    int tile_shift = 0;
    if (ctx->opt_persistent && (variant == 1 || variant == 65 || variant == 129 || variant == 193)) {
        variant |= 256;
        if (ctx->opt_persistent == 2 && variant == 449) variant |= 512;
        const unsigned fill = (unsigned)(ctx->num_sms * ((variant & 512) ? 10 : 12) * (TRACE_THREADS / TB));      // blocks that are resident at once
        if (grid > fill) grid = fill;
        tile_shift = TB == 128 ? 2 : TB == 64 ? 1 : 0;
        cudaError_t me = cudaMemsetAsync(ctx->d_counters + 4, 0, sizeof(unsigned long long), stream);      // the work ticket
        if (me != cudaSuccess) return lrc_fail(ctx, LRC_ERR_CUDA, "cudaMemsetAsync(ticket): %s", cudaGetErrorString(me));
    }
    const int top_n = (variant & 8) ? (int)((1 << ctx->opt_top_levels) - 1) : 0;
    const int stack_levels = (variant & 16) ? (int)ctx->opt_stack_levels : 0;
    cfg.dynamicSmemBytes = (variant & 16) ? (size_t)stack_levels * LRC_SS_STRIDE * sizeof(int) : (size_t)top_n * 64;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3((unsigned)TB); cfg.stream = stream;
    cudaError_t le = cudaSuccess;
#define LRC_LAUNCH_TRACE(COUNT, VARIANT) \
    le = cudaLaunchKernelEx(&cfg, k_trace<MODE, COUNT, DENSE, VARIANT>, g, fm, nodes, tris, n, has_tris, hp, inc, block_count, t_hit, prim, counters, top_table, top_n, stack_levels, nq, root, tile_shift, ntex)
    if (ctx->counting) {
        switch (variant) {
            case 0: LRC_LAUNCH_TRACE(true, 0); break;
            case 1: LRC_LAUNCH_TRACE(true, 1); break;
            case 2: LRC_LAUNCH_TRACE(true, 2); break;
            case 5: LRC_LAUNCH_TRACE(true, 5); break;
            case 13: LRC_LAUNCH_TRACE(true, 13); break;
            case 21: LRC_LAUNCH_TRACE(true, 21); break;
            case 37: LRC_LAUNCH_TRACE(true, 37); break;
            case 65: LRC_LAUNCH_TRACE(true, 65); break;
            case 129: LRC_LAUNCH_TRACE(true, 129); break;
            case 193: LRC_LAUNCH_TRACE(true, 193); break;
            case 257: LRC_LAUNCH_TRACE(true, 257); break;
            case 321: LRC_LAUNCH_TRACE(true, 321); break;
            case 385: LRC_LAUNCH_TRACE(true, 385); break;
            case 449: LRC_LAUNCH_TRACE(true, 449); break;
            case 961: LRC_LAUNCH_TRACE(true, 961); break;
            case 1217: LRC_LAUNCH_TRACE(true, 1217); break;
            case 2241: LRC_LAUNCH_TRACE(true, 2241); break;
            case 4289: LRC_LAUNCH_TRACE(true, 4289); break;
            case 6337: LRC_LAUNCH_TRACE(true, 6337); break;
            case 7361: LRC_LAUNCH_TRACE(true, 7361); break;
            case 10433: LRC_LAUNCH_TRACE(true, 10433); break;
            case 18625: LRC_LAUNCH_TRACE(true, 18625); break;
            case 26817: LRC_LAUNCH_TRACE(true, 26817); break;
            default: LRC_LAUNCH_TRACE(true, 3); break;
        }
    } else {
        switch (variant) {
            case 0: LRC_LAUNCH_TRACE(false, 0); break;
            case 1: LRC_LAUNCH_TRACE(false, 1); break;
            case 2: LRC_LAUNCH_TRACE(false, 2); break;
            case 5: LRC_LAUNCH_TRACE(false, 5); break;
            case 13: LRC_LAUNCH_TRACE(false, 13); break;
            case 21: LRC_LAUNCH_TRACE(false, 21); break;
            case 37: LRC_LAUNCH_TRACE(false, 37); break;
            case 65: LRC_LAUNCH_TRACE(false, 65); break;
            case 129: LRC_LAUNCH_TRACE(false, 129); break;
            case 193: LRC_LAUNCH_TRACE(false, 193); break;
            case 257: LRC_LAUNCH_TRACE(false, 257); break;
            case 321: LRC_LAUNCH_TRACE(false, 321); break;
            case 385: LRC_LAUNCH_TRACE(false, 385); break;
            case 449: LRC_LAUNCH_TRACE(false, 449); break;
            case 961: LRC_LAUNCH_TRACE(false, 961); break;
            case 1217: LRC_LAUNCH_TRACE(false, 1217); break;
            case 2241: LRC_LAUNCH_TRACE(false, 2241); break;
            case 4289: LRC_LAUNCH_TRACE(false, 4289); break;
            case 6337: LRC_LAUNCH_TRACE(false, 6337); break;
            case 7361: LRC_LAUNCH_TRACE(false, 7361); break;
            case 10433: LRC_LAUNCH_TRACE(false, 10433); break;
            case 18625: LRC_LAUNCH_TRACE(false, 18625); break;
            case 26817: LRC_LAUNCH_TRACE(false, 26817); break;
            default: LRC_LAUNCH_TRACE(false, 3); break;
        }
    }
#undef LRC_LAUNCH_TRACE
    if (le != cudaSuccess) return lrc_fail(ctx, LRC_ERR_CUDA, "launch k_trace: %s", cudaGetErrorString(le));
    LRC_CHECK_LAUNCH(ctx, "k_trace");
    return LRC_OK;
}

int check_out(lrc_ctx* ctx, const lrc_out* out, int64_t need)
{
    if (!out || !out->xyz || !out->frame_offset) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_out: xyz and frame_offset are required");
    if (out->capacity < need) return lrc_fail(ctx, LRC_ERR_CAPACITY, "lrc_out: capacity is smaller than the number of rays");
    return LRC_OK;
}

// Shared driver of all scan-type entry points: per chunk of whole frames, trace (+ block counts) -> scan of the
// counts -> streaming compaction, with a running output offset carried in device memory (no host round trip).
// With more than one chunk the compaction of chunk c (and, when gather targets are set, its NVLink stores) runs on
// an auxiliary stream while chunk c+1 is traversed on the caller's stream; the scratch is double-buffered.
template <int MODE>
int run_scan(lrc_ctx* ctx, RayGen g, int64_t P, int64_t N, const double* h_center, double max_range, lrc_out* out,
             cudaStream_t stream)
{
    const int64_t total = P * N;
    int rc = check_out(ctx, out, total);
    if (rc) return rc;
    if ((rc = ensure_counters(ctx))) return rc;
    const bool gather = ctx->gather.n > 0;
    if (gather && ctx->gather.capacity < total) return lrc_fail(ctx, LRC_ERR_CAPACITY, "lrc_set_gather: capacity is smaller than the number of rays");
    if (gather && ctx->gather.frame_capacity > 0 && P + 1 > ctx->gather.frame_capacity)
        return lrc_fail(ctx, LRC_ERR_CAPACITY, "lrc_set_gather: frame_capacity is smaller than the number of frames + 1");
    if (total == 0) {
        LRC_CUDA(ctx, cudaMemsetAsync(out->frame_offset, 0, sizeof(int64_t) * (P + 1), stream));
        return LRC_OK;
    }
    // chunk by whole frames: bounded scratch (24 B per ray per slot) and, with gather targets, overlap
    int64_t frames_per_chunk = ctx->opt_chunk_rays / N;
    if (frames_per_chunk < 1) frames_per_chunk = 1;
    if (frames_per_chunk > P) frames_per_chunk = P;   // size the scratch by the largest REAL chunk (a one-frame call must not reserve 2^26 rays)
    if (MODE == MODE_RAYS) frames_per_chunk = P;   // a single explicit frame
    std::vector<int64_t> chunk_start;              // first frame of every chunk, closed by P
    // Pose chunks: chunk c is compacted (and, with gather targets, exchanged) on the auxiliary stream while chunk c+1 is
    // traversed.  Without gather targets ("scan_chunks" / "scan_taper") this hides the HBM-bound compaction behind k_trace,
    // which leaves DRAM idle; only a whole trajectory of at least ~1M rays per chunk is cut.
    int64_t want_chunks = 1, ramp = 1, taper = 1;
    if (MODE != MODE_RAYS && P > 1) {
        if (gather) { want_chunks = ctx->opt_gather_chunks; ramp = ctx->opt_gather_ramp; taper = ctx->opt_gather_taper; }
        else {
            want_chunks = ctx->opt_scan_chunks; taper = ctx->opt_scan_taper;
            const int64_t by_size = total / ((int64_t)1 << 20);
            if (want_chunks > by_size) want_chunks = by_size;
        }
    }
    if (want_chunks > 1) {
        // With gather targets two things are exposed: the wait for the first chunk (the exchange, not the traversal, is the
        // long pole at 4+ GPUs: NVLink ingress) and the exchange of the last chunk.  "gather_ramp" / "gather_taper" make the
        // first / last chunk 1/ramp / 1/taper of a regular one.
        const int64_t C = want_chunks < P ? want_chunks : P;
        std::vector<double> wgt((size_t)C, 1.0);
        if (C > 1) { wgt[0] = 1.0 / (double)ramp; wgt[(size_t)C - 1] = 1.0 / (double)taper; }
        double sum = 0.0;
        for (double x : wgt) sum += x;
        double acc = 0.0;
        int64_t f = 0;
        for (int64_t c = 0; c < C && f < P; ++c) {
            chunk_start.push_back(f);
            acc += wgt[(size_t)c];
            int64_t end = c == C - 1 ? P : (int64_t)((double)P * acc / sum + 0.5);
            if (end <= f) end = f + 1;
            if (end - f > frames_per_chunk) end = f + frames_per_chunk;     // never beyond the scratch bound
            if (end > P) end = P;
            f = end;
        }
        while (f < P) { chunk_start.push_back(f); f += frames_per_chunk < P - f ? frames_per_chunk : P - f; }
    } else {
        for (int64_t f = 0; f < P; f += frames_per_chunk) chunk_start.push_back(f);
    }
    chunk_start.push_back(P);
    const int64_t n_chunks = (int64_t)chunk_start.size() - 1;
    frames_per_chunk = 1;
    for (int64_t c = 0; c < n_chunks; ++c)
        if (chunk_start[(size_t)c + 1] - chunk_start[(size_t)c] > frames_per_chunk) frames_per_chunk = chunk_start[(size_t)c + 1] - chunk_start[(size_t)c];
    const bool piped = n_chunks > 1;
    const int n_slots = piped ? 2 : 1;
    const int64_t chunk_rays = frames_per_chunk * N;
    const int TB = scan_block_threads(ctx, total);
    ctx->cur_block = TB;
    const int64_t max_blocks = (chunk_rays + TB - 1) / TB;
    const bool wire = gather && ctx->wire.enabled && MODE != MODE_RAYS;
    if (wire && out->incident_deg)
        return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_gather_wire: the scan's lrc_out must not carry incident_deg (use lrc_incident_angles)");
    if (wire && !out->label && ctx->T > 0) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_gather_wire: the scan's lrc_out needs a label array");
    const bool want_inc = (out->incident_deg != nullptr && max_range >= 0.0) || wire;      // wire: the slot carries t
    const size_t hp_bytes = align_up(sizeof(float4) * (size_t)chunk_rays, 256);
    const size_t inc_bytes = want_inc ? align_up(sizeof(double) * (size_t)chunk_rays, 256) : 0;
    const size_t slot_bytes = hp_bytes + inc_bytes;
    if ((rc = lrc_grow(ctx, &ctx->scratch, &ctx->scratch_bytes, slot_bytes * n_slots))) return rc;
    // the scratch is shared by every scan of this context: order this call after the previous one, whatever its stream
    if (!ctx->scratch_event) LRC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->scratch_event, cudaEventDisableTiming));
    else LRC_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->scratch_event, 0));
    const size_t cnt_bytes = align_up(sizeof(unsigned) * (size_t)max_blocks, 256);
    const size_t need2 = 2 * cnt_bytes * n_slots + sizeof(long long) * (size_t)(n_chunks + 1);
    if ((rc = lrc_grow(ctx, &ctx->scratch2, &ctx->scratch2_bytes, need2))) return rc;
    char* b2 = (char*)ctx->scratch2;
    long long* run = (long long*)(b2 + 2 * cnt_bytes * n_slots);
    LRC_CUDA(ctx, cudaMemsetAsync(run, 0, sizeof(long long), stream));
    cudaStream_t aux = stream;
    if (piped) {
        if (!ctx->s_aux) {
            int lo = 0, hi = 0;
            LRC_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
            LRC_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->s_aux, cudaStreamNonBlocking, hi));   // compaction blocks are short: let them in first
            for (int k = 0; k < 4; ++k) LRC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->pipe_ev[k], cudaEventDisableTiming));
        }
        aux = ctx->s_aux;
        LRC_CUDA(ctx, cudaEventRecord(ctx->pipe_ev[3], stream));      // aux must see the memset of run[0] and the tables
        LRC_CUDA(ctx, cudaStreamWaitEvent(aux, ctx->pipe_ev[3], 0));
    }
    const bool timing = ctx->opt_kernel_timing != 0;
    if (timing) {
        ctx->kt_used = 0;
        while (ctx->kt_events.size() < (size_t)(4 * n_chunks)) {
            cudaEvent_t e;
            LRC_CUDA(ctx, cudaEventCreate(&e));
            ctx->kt_events.push_back(e);
        }
    }
    FrameMath fm;
    fm.max_range = max_range;
    fm.cx = h_center ? h_center[0] : 0.0; fm.cy = h_center ? h_center[1] : 0.0; fm.cz = h_center ? h_center[2] : 0.0;
    fm.wire = wire ? 1 : 0;
    GatherTargets gt;
    memset(&gt, 0, sizeof gt);
    long long wire_tag = 0;
    if (wire) {
        const lrc_gather_wire& W = ctx->wire;
        if (W.rank_frames[W.self] != P)
            return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_gather_wire: this scan's frame count differs from rank_frames[self]");
        if (!ctx->s_rebuild) {
            LRC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_rebuild, cudaStreamNonBlocking));
            LRC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->rebuild_ev, cudaEventDisableTiming));
            LRC_CUDA(ctx, cudaMalloc((void**)&ctx->d_wire_err, sizeof(int)));
            LRC_CUDA(ctx, cudaMemset(ctx->d_wire_err, 0, sizeof(int)));
        }
        wire_tag = (long long)(++ctx->wire_scan) << 32;
        // the rebuild stream reads the ray tables of this call and must not run ahead of it
        LRC_CUDA(ctx, cudaEventRecord(ctx->rebuild_ev, stream));
        LRC_CUDA(ctx, cudaStreamWaitEvent(ctx->s_rebuild, ctx->rebuild_ev, 0));
    }
    if (gather) {
        gt.wire = wire ? 1 : 0;
        gt.self = wire ? ctx->wire.self : 0;
        if (wire)
            for (int k = 0; k < ctx->gather.n; ++k) { gt.wire_t[k] = ctx->wire.t[k]; gt.wire_ray[k] = ctx->wire.ray_idx[k]; gt.ready[k] = ctx->wire.ready[k]; }
        gt.n = ctx->gather.n;
        for (int k = 0; k < gt.n; ++k) { gt.xyz[k] = ctx->gather.xyz[k]; gt.label[k] = ctx->gather.label[k]; gt.frame_offset[k] = ctx->gather.frame_offset[k]; }
        gt.point_base = ctx->gather.point_base; gt.frame_base = ctx->gather.frame_base; gt.capacity = ctx->gather.capacity;
    }
    for (int64_t c = 0; c < n_chunks; ++c) {
        const int slot = piped ? (int)(c & 1) : 0;
        const int64_t f0 = chunk_start[(size_t)c];
        const int64_t nf = chunk_start[(size_t)c + 1] - f0;
        const int64_t n = nf * N;
        const int64_t nb = (n + TB - 1) / TB;
        float4* hp = (float4*)((char*)ctx->scratch + slot_bytes * slot);
        double* inc = want_inc ? (double*)((char*)ctx->scratch + slot_bytes * slot + hp_bytes) : nullptr;
        unsigned* counts = (unsigned*)(b2 + 2 * cnt_bytes * slot);
        unsigned* base = (unsigned*)(b2 + 2 * cnt_bytes * slot + cnt_bytes);
        g.pose0 = f0;
        if (piped && c >= 2) LRC_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->pipe_ev[slot], 0));   // slot free again?
        if (timing) LRC_CUDA(ctx, cudaEventRecord(ctx->kt_events[4 * c + 0], stream));
        if ((rc = launch_trace<MODE, false>(ctx, g, fm, n, hp, inc, counts, nullptr, nullptr, stream))) return rc;
        if (timing) LRC_CUDA(ctx, cudaEventRecord(ctx->kt_events[4 * c + 1], stream));
        if (piped) {
            LRC_CUDA(ctx, cudaEventRecord(ctx->pipe_ev[2], stream));
            LRC_CUDA(ctx, cudaStreamWaitEvent(aux, ctx->pipe_ev[2], 0));
        }
        const bool last = (c == n_chunks - 1);
        if (timing) LRC_CUDA(ctx, cudaEventRecord(ctx->kt_events[4 * c + 2], aux));
        k_scan_counts<<<1, 1024, 0, aux>>>(counts, base, nb, run + c, run + c + 1, last ? out->frame_offset + P : nullptr);
        LRC_CHECK_LAUNCH(ctx, "k_scan_counts");
        CompactParams q;
        q.hp = hp; q.inc = inc; q.base = base; q.run_in = run + c;
        q.n = n; q.ray0 = f0 * N; q.N = N; q.total_rays = total; q.P = P;
        q.labels = ctx->T > 0 ? ctx->labels : nullptr;
        q.out = *out;
        q.wire_t = wire ? ctx->wire.t[ctx->wire.self] + ctx->gather.point_base : nullptr;
        q.wire_ray = wire ? ctx->wire.ray_idx[ctx->wire.self] + ctx->gather.point_base : nullptr;
        if (!want_inc || wire) q.out.incident_deg = nullptr;
        if (ctx->T == 0) q.out.label = nullptr;
        if (gather && !q.out.label && ctx->T > 0) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_gather: the scan's lrc_out needs a label array");
        k_compact<<<(unsigned)nb, TB, 0, aux>>>(q);
        LRC_CHECK_LAUNCH(ctx, "k_compact");
        // the scratch slot is free as soon as the compaction has read it: the traversal of chunk c+2 must not wait for
        // the exchange of chunk c
        if (piped) LRC_CUDA(ctx, cudaEventRecord(ctx->pipe_ev[slot], aux));
        if (gather) {
            PushParams pp;
            pp.xyz = out->xyz; pp.label = q.out.label; pp.frame_offset = out->frame_offset; pp.run = run + c;
            pp.f0 = f0; pp.nf = nf; pp.P = P; pp.last = last ? 1 : 0;
            if (ctx->opt_push_mode == 1 || wire) {
                if (!ctx->push_tma_ready) {      // per context (= per device): the opt-in to 64 KB of dynamic shared memory
                    LRC_CUDA(ctx, cudaFuncSetAttribute(k_push_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_STAGES * TMA_TILE_MAX));
                    ctx->push_tma_ready = true;
                }
                k_push_tma<<<(unsigned)ctx->opt_push_blocks, TMA_THREADS, TMA_STAGES * (size_t)ctx->opt_push_tile, aux>>>(pp, gt, (int)ctx->opt_push_tile);
                LRC_CHECK_LAUNCH(ctx, "k_push_tma");
            } else {
                k_push<<<dim3((unsigned)ctx->opt_push_blocks, (unsigned)gt.n), PUSH_THREADS, 0, aux>>>(pp, gt);
                LRC_CHECK_LAUNCH(ctx, "k_push");
            }
            if (wire) {
                // tell every target that frames [0, f0 + nf) of this scan are complete there ...
                k_wire_flag<<<1, 32, 0, aux>>>(gt, wire_tag | (long long)(f0 + nf));
                LRC_CHECK_LAUNCH(ctx, "k_wire_flag");
                // ... and rebuild the points of the same frame range of every other rank as soon as THEIR words arrive
                const lrc_gather_wire& W = ctx->wire;
                WireWait ww;
                memset(&ww, 0, sizeof ww);
                ww.ready = W.ready[W.self];
                ww.n = gt.n;
                RebuildParams rp;
                memset(&rp, 0, sizeof rp);
                rp.xyz = ctx->gather.xyz[W.self]; rp.wire_t = W.t[W.self]; rp.wire_ray = W.ray_idx[W.self];
                rp.frame_offset = ctx->gather.frame_offset[W.self];
                rp.n = gt.n; rp.self = W.self;
                rp.pose_index_base = g.pose_index_base - (uint64_t)W.rank_pose0[W.self];
                bool any = false;
                for (int p = 0; p < gt.n; ++p) {
                    const int64_t Pp = W.rank_frames[p];
                    rp.point_base[p] = W.rank_point_base[p]; rp.frame_base[p] = W.rank_frame_base[p]; rp.pose0[p] = W.rank_pose0[p];
                    rp.fa[p] = f0 < Pp ? f0 : Pp;
                    rp.fb[p] = last ? Pp : (f0 + nf < Pp ? f0 + nf : Pp);
                    if (p != W.self && rp.fb[p] > rp.fa[p]) { ww.need[p] = wire_tag | (long long)rp.fb[p]; any = true; }
                }
                if (any) {
                    k_wire_wait<<<1, 32, 0, ctx->s_rebuild>>>(ww, ctx->d_wire_err);
                    LRC_CHECK_LAUNCH(ctx, "k_wire_wait");
                    RayGen ga = g;
                    ga.poses = W.all_poses;
                    int64_t max_frames = 0;
                    for (int p = 0; p < gt.n; ++p)
                        if (p != W.self && rp.fb[p] - rp.fa[p] > max_frames) max_frames = rp.fb[p] - rp.fa[p];
                    int sub_blocks = (int)((N + 256 * 8 - 1) / (256 * 8));        // ~8 points per thread
                    if (sub_blocks < 1) sub_blocks = 1;
                    k_wire_rebuild<MODE><<<dim3((unsigned)(max_frames * sub_blocks), (unsigned)gt.n), 256, 0, ctx->s_rebuild>>>(ga, rp, sub_blocks);
                    LRC_CHECK_LAUNCH(ctx, "k_wire_rebuild");
                }
            }
        }
        if (timing) { LRC_CUDA(ctx, cudaEventRecord(ctx->kt_events[4 * c + 3], aux)); ctx->kt_used = (size_t)(4 * (c + 1)); }
    }
    if (gather && piped) {   // the last exchange kernel on the auxiliary stream
        LRC_CUDA(ctx, cudaEventRecord(ctx->pipe_ev[2], aux));
        LRC_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->pipe_ev[2], 0));
    }
    if (piped) {   // the caller's stream owns the result: join the auxiliary stream back
        const int last_slot = (int)((n_chunks - 1) & 1);
        LRC_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->pipe_ev[last_slot], 0));
        if (n_chunks >= 2) LRC_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->pipe_ev[last_slot ^ 1], 0));
    }
    if (wire) {      // the caller's stream owns the whole gathered cloud, rebuilt points included
        LRC_CUDA(ctx, cudaEventRecord(ctx->rebuild_ev, ctx->s_rebuild));
        LRC_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->rebuild_ev, 0));
    }
    LRC_CUDA(ctx, cudaEventRecord(ctx->scratch_event, stream));
    if (out->incident_deg && !want_inc)   // rays_intersect_mesh flavour: no angles are defined; keep the buffer deterministic
        LRC_CUDA(ctx, cudaMemsetAsync(out->incident_deg, 0, sizeof(double) * (size_t)(total < out->capacity ? total : out->capacity), stream));
    return LRC_OK;
}

int fill_single(lrc_ctx* ctx, const lrc_single_axis* s, const double* poses, RayGen& g, cudaStream_t stream)
{
    if (!s || s->H < 1 || s->W < 1 || s->H > LRC_MAX_H) return lrc_fail(ctx, LRC_ERR_INVALID, "single-axis sensor: need 1 <= H <= 4096 and W >= 1");
    if ((int64_t)s->H * s->W >= (int64_t)1 << 31) return lrc_fail(ctx, LRC_ERR_INVALID, "single-axis sensor: H*W must be < 2^31");
    memset(&g, 0, sizeof g);
    g.poses = poses;
    g.H = s->H; g.W = s->W; g.N = s->H * s->W;
    g.uniform_mode = s->h_vertical_deg ? 0 : 1;
    const size_t n_tab = (size_t)2 * s->W + 2 * s->H;
    // the per-waypoint call pattern presents the same sensor thousands of times: keep the tables while the sensor is unchanged
    std::vector<double> key;
    key.reserve(5 + (size_t)(s->h_vertical_deg ? s->H : 0));
    key.push_back((double)s->W); key.push_back((double)s->H); key.push_back((double)g.uniform_mode);
    key.push_back(s->fov_up_deg); key.push_back(s->fov_down_deg);
    if (s->h_vertical_deg) key.insert(key.end(), s->h_vertical_deg, s->h_vertical_deg + s->H);
    const bool same = ctx->tables && key.size() == ctx->tables_key.size() &&
                      memcmp(key.data(), ctx->tables_key.data(), sizeof(double) * key.size()) == 0;
    if (!same) {
        ctx->tables_key.clear();
        int rc = lrc_grow(ctx, (void**)&ctx->tables, &ctx->tables_bytes, sizeof(double) * (n_tab + s->H));
        if (rc) return rc;
        double* vdeg = ctx->tables + n_tab;
        // the tables are shared by every scan of this context: a scan still running on another stream must finish with them first
        if (ctx->scratch_event) LRC_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->scratch_event, 0));
        if (s->h_vertical_deg)
            LRC_CUDA(ctx, cudaMemcpyAsync(vdeg, s->h_vertical_deg, sizeof(double) * s->H, cudaMemcpyHostToDevice, stream));
        const int n = s->W + s->H;
        k_single_tables<<<(n + 255) / 256, 256, 0, stream>>>(ctx->tables, s->W, s->H, vdeg, g.uniform_mode, s->fov_up_deg, s->fov_down_deg);
        LRC_CHECK_LAUNCH(ctx, "k_single_tables");
        // later scans may run on other streams: they must see the finished tables
        if (!ctx->tables_event) LRC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->tables_event, cudaEventDisableTiming));
        LRC_CUDA(ctx, cudaEventRecord(ctx->tables_event, stream));
        ctx->tables_key.swap(key);
    } else if (ctx->tables_event) {
        LRC_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->tables_event, 0));
    }
    g.tab = ctx->tables;
    return LRC_OK;
}

int fill_dual(lrc_ctx* ctx, const lrc_dual_axis* s, const double* poses, RayGen& g)
{
    if (!s || s->num_lines < 1 || s->points_per_line < 1) return lrc_fail(ctx, LRC_ERR_INVALID, "dual-axis sensor: need num_lines >= 1 and points_per_line >= 1");
    if ((int64_t)s->num_lines * s->points_per_line >= (int64_t)1 << 31) return lrc_fail(ctx, LRC_ERR_INVALID, "dual-axis sensor: too many rays per frame");
    memset(&g, 0, sizeof g);
    g.poses = poses;
    g.H = s->num_lines; g.W = s->points_per_line; g.N = s->num_lines * s->points_per_line;
    g.theta_min = s->theta_min; g.theta_max = s->theta_max;
    g.tstep = s->num_lines > 1 ? (s->theta_min - s->theta_max) / (double)(s->num_lines - 1) : 0.0;   // linspace(theta_max, theta_min, L)
    g.pstep = (2.0 * PI_D - 0.0) / (double)s->points_per_line;                                         // linspace(0, 2pi, K, endpoint=False)
    g.swing_amp = s->swing_amplitude; g.swing_freq = s->swing_frequency;
    return LRC_OK;
}

void fill_noise(const lrc_noise* nz, RayGen& g, bool dual)
{
    if (!nz) return;
    if (dual) { g.angle_std = nz->angle_noise_std; g.dropout_p = nz->dropout_probability; }
    g.range_std = nz->range_noise_std;
    g.k0 = (uint32_t)nz->seed; g.k1 = (uint32_t)(nz->seed >> 32);
    g.pose_index_base = nz->pose_index_base;
}

int precheck(lrc_ctx* ctx, const char* who)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "%s: ctx is NULL", who);
    if (!ctx->has_mesh) return lrc_fail(ctx, LRC_ERR_NO_MESH, "%s: call lrc_set_mesh first", who);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return lrc_fail(ctx, LRC_ERR_CUDA, "%s: cudaSetDevice: %s", who, cudaGetErrorString(e));
    return LRC_OK;
}

}  // namespace

// ======================================================================================================
extern "C" int lrc_abi_version(void) { return LRC_ABI_VERSION; }

#define LRC_DEFAULT_L2_PERSIST 0
extern "C" int lrc_default_l2_persist(void) { return LRC_DEFAULT_L2_PERSIST; }

extern "C" int lrc_create(int device, lrc_ctx** out)
{
    if (!out) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        return lrc_fail(nullptr, LRC_ERR_NO_DEVICE, "lrc_create: no CUDA device (%s); this engine has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_create: device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return lrc_fail(nullptr, LRC_ERR_CUDA, "lrc_create: cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return lrc_fail(nullptr, LRC_ERR_NO_DEVICE, "lrc_create: device '%s' is not sm_100 (this library is built for B200 only)", prop.name);
    if ((e = cudaSetDevice(device)) != cudaSuccess)
        return lrc_fail(nullptr, LRC_ERR_CUDA, "lrc_create: cudaSetDevice: %s", cudaGetErrorString(e));
    lrc_ctx* ctx = new lrc_ctx();
    ctx->device = device;
    ctx->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    ctx->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    ctx->l2_size = (size_t)prop.l2CacheSize;
    ctx->num_sms = prop.multiProcessorCount;
    *out = ctx;
    return LRC_OK;
}

extern "C" void lrc_destroy(lrc_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->nodes_tex) cudaDestroyTextureObject(ctx->nodes_tex);
    cudaFree(ctx->bvh_block); cudaFree(ctx->labels); cudaFree(ctx->top_table);
    cudaFree(ctx->scratch); cudaFree(ctx->scratch2); cudaFree(ctx->tables); cudaFree(ctx->d_counters);
    cudaFree(ctx->host_dev); cudaFree(ctx->mesh_dev); cudaFree(ctx->post_scratch);
    cudaFree(ctx->ci_meta); cudaFree(ctx->ci_start); cudaFree(ctx->ci_sorted); cudaFree(ctx->cg_scratch);
    cudaFree(ctx->nn_meta); cudaFree(ctx->nn_start); cudaFree(ctx->nn_sorted);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    if (ctx->scratch_event) cudaEventDestroy(ctx->scratch_event);
    if (ctx->tables_event) cudaEventDestroy(ctx->tables_event);
    if (ctx->s_rebuild) { cudaStreamDestroy(ctx->s_rebuild); cudaEventDestroy(ctx->rebuild_ev); cudaFree(ctx->d_wire_err); }
    for (cudaEvent_t e : ctx->kt_events) cudaEventDestroy(e);
    if (ctx->s_aux) { cudaStreamDestroy(ctx->s_aux); for (int k = 0; k < 4; ++k) cudaEventDestroy(ctx->pipe_ev[k]); }
    for (size_t i = 0; i < ctx->n_events; ++i) cudaEventDestroy(ctx->events[i]);
    free(ctx->events);
    if (ctx->s_compute) cudaStreamDestroy(ctx->s_compute);
    if (ctx->s_copy) cudaStreamDestroy(ctx->s_copy);
    if (ctx->s_count) cudaStreamDestroy(ctx->s_count);
    delete ctx;
}

extern "C" const char* lrc_last_error(const lrc_ctx* ctx) { return ctx ? ctx->err : g_lrc_global_err; }

extern "C" int64_t lrc_launch_count(const lrc_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int lrc_set_counting(lrc_ctx* ctx, int enabled)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_set_counting: ctx is NULL");
    ctx->counting = enabled ? 1 : 0;
    return ensure_counters(ctx);
}

extern "C" int lrc_counters(lrc_ctx* ctx, lrc_counters_t* h_out, int reset, void* stream_)
{
    if (!ctx || !h_out) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_counters: NULL argument");
    int rc = ensure_counters(ctx);
    if (rc) return rc;
    cudaStream_t stream = (cudaStream_t)stream_;
    unsigned long long h[4];
    LRC_CUDA(ctx, cudaMemcpyAsync(h, ctx->d_counters, sizeof h, cudaMemcpyDeviceToHost, stream));
    if (reset) LRC_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, sizeof h, stream));
    LRC_CUDA(ctx, cudaStreamSynchronize(stream));
    h_out->rays = h[0]; h_out->nodes_visited = h[1]; h_out->tris_tested = h[2]; h_out->hits = h[3];
    return LRC_OK;
}

extern "C" int lrc_get_stat(lrc_ctx* ctx, const char* key, int64_t* h_value)
{
    if (!ctx || !key || !h_value) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_get_stat: NULL argument");
    if (!strcmp(key, "scratch_bytes")) { *h_value = (int64_t)(ctx->scratch_bytes + ctx->scratch2_bytes); return LRC_OK; }
    if (!strcmp(key, "bvh_bytes")) { *h_value = (int64_t)ctx->bvh_bytes; return LRC_OK; }
    if (!strcmp(key, "build_quality")) { *h_value = ctx->build_quality; return LRC_OK; }
    if (!strcmp(key, "ploc_iterations")) { *h_value = ctx->ploc_iterations; return LRC_OK; }
    if (!strcmp(key, "node_format")) { *h_value = ctx->node_format; return LRC_OK; }
    if (!strcmp(key, "leaf_size")) { *h_value = ctx->opt_leaf_size; return LRC_OK; }
    if (!strcmp(key, "variant")) { *h_value = ctx->opt_variant; return LRC_OK; }
    if (!strcmp(key, "tune")) { *h_value = ctx->opt_tune; return LRC_OK; }
    if (!strcmp(key, "warp_packet")) { *h_value = ctx->opt_warp_packet; return LRC_OK; }
    if (!strcmp(key, "rays_per_thread")) { *h_value = ctx->opt_rays_per_thread; return LRC_OK; }
    if (!strcmp(key, "persistent")) { *h_value = ctx->opt_persistent; return LRC_OK; }
    if (!strcmp(key, "num_sms")) { *h_value = ctx->num_sms; return LRC_OK; }
    if (!strcmp(key, "wire_error")) {      // 1: a peer's progress word did not arrive in time during some scan (synchronises the device)
        int e = 0;
        if (ctx->d_wire_err) {
            LRC_CUDA(ctx, cudaSetDevice(ctx->device));
            LRC_CUDA(ctx, cudaDeviceSynchronize());
            LRC_CUDA(ctx, cudaMemcpy(&e, ctx->d_wire_err, sizeof e, cudaMemcpyDeviceToHost));
        }
        *h_value = e;
        return LRC_OK;
    }
    if (!strcmp(key, "nn_generation")) { *h_value = ctx->nn_generation; return LRC_OK; }
    if (!strcmp(key, "collision_generation")) { *h_value = ctx->ci_generation; return LRC_OK; }
    if (!strcmp(key, "mesh_generation")) { *h_value = ctx->mesh_generation; return LRC_OK; }
    return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_get_stat: unknown key '%s'", key);
}

extern "C" int lrc_kernel_times(lrc_ctx* ctx, double* h_trace_ms, double* h_compact_ms, int32_t* h_launches)
{
    if (!ctx || !h_trace_ms || !h_compact_ms) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_kernel_times: NULL argument");
    if (!ctx->opt_kernel_timing || ctx->kt_used == 0)
        return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_kernel_times: enable option kernel_timing and run a scan first");
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    double tr = 0.0, cp = 0.0;
    for (size_t c = 0; c < ctx->kt_used / 4; ++c) {
        float a = 0.f, b = 0.f;
        LRC_CUDA(ctx, cudaEventSynchronize(ctx->kt_events[4 * c + 1]));
        LRC_CUDA(ctx, cudaEventSynchronize(ctx->kt_events[4 * c + 3]));
        LRC_CUDA(ctx, cudaEventElapsedTime(&a, ctx->kt_events[4 * c + 0], ctx->kt_events[4 * c + 1]));
        LRC_CUDA(ctx, cudaEventElapsedTime(&b, ctx->kt_events[4 * c + 2], ctx->kt_events[4 * c + 3]));
        tr += a; cp += b;
    }
    *h_trace_ms = tr;
    *h_compact_ms = cp;
    if (h_launches) *h_launches = (int32_t)(ctx->kt_used / 4);
    return LRC_OK;
}

extern "C" int lrc_set_option(lrc_ctx* ctx, const char* key, int64_t value)
{
    if (!ctx || !key) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_option: NULL argument");
    if (!strcmp(key, "chunk_rays")) { if (value < 1) return lrc_fail(ctx, LRC_ERR_INVALID, "chunk_rays must be >= 1"); ctx->opt_chunk_rays = value; return LRC_OK; }
    if (!strcmp(key, "l2_persist")) {
        // value = percentage of the device's maximum persisting-L2 set-aside to reserve (0 = off)
        if (value < 0 || value > 100) return lrc_fail(ctx, LRC_ERR_INVALID, "l2_persist must be in [0, 100]");
        LRC_CUDA(ctx, cudaSetDevice(ctx->device));
        cudaDeviceProp prop;
        LRC_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
        const size_t want = (size_t)((double)prop.persistingL2CacheMaxSize * (double)value / 100.0);
        LRC_CUDA(ctx, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
        if (value == 0) LRC_CUDA(ctx, cudaCtxResetPersistingL2Cache());
        ctx->l2_persist_max = want;
        ctx->opt_l2_persist = value;
        return LRC_OK;
    }
    if (!strcmp(key, "l2_reset")) { LRC_CUDA(ctx, cudaSetDevice(ctx->device)); LRC_CUDA(ctx, cudaCtxResetPersistingL2Cache()); return LRC_OK; }
    if (!strcmp(key, "leaf_size")) {
        if (value < 1 || value > 8) return lrc_fail(ctx, LRC_ERR_INVALID, "leaf_size must be in [1, 8]");
        ctx->opt_leaf_size = value;        // takes effect at the next lrc_set_mesh
        return LRC_OK;
    }
    if (!strcmp(key, "build_quality")) {
        if (value != 0 && value != 1) return lrc_fail(ctx, LRC_ERR_INVALID, "build_quality must be 0 (LBVH) or 1 (PLOC)");
        ctx->opt_build_quality = value;    // takes effect at the next lrc_set_mesh
        return LRC_OK;
    }
    if (!strcmp(key, "compact_nodes")) { ctx->opt_compact_nodes = value != 0; return LRC_OK; }      // takes effect at the next lrc_set_mesh
    if (!strcmp(key, "ploc_radius")) {
        if (value < 1 || value > 32) return lrc_fail(ctx, LRC_ERR_INVALID, "ploc_radius must be in [1, 32]");
        ctx->opt_ploc_radius = value;
        return LRC_OK;
    }
    if (!strcmp(key, "node_format")) {
        if (value < 0 || value > 2) return lrc_fail(ctx, LRC_ERR_INVALID, "node_format must be 0 (64 B float boxes), 1 (32 B 16-bit boxes) or 2 (64 B paired boxes, packed FMA)");
        ctx->opt_node_format = value;      // takes effect at the next lrc_set_mesh
        return LRC_OK;
    }
    if (!strcmp(key, "stack_levels")) {
        if (value < 1 || value > 48) return lrc_fail(ctx, LRC_ERR_INVALID, "stack_levels must be in [1, 48]");
        ctx->opt_stack_levels = value;
        return LRC_OK;
    }
    if (!strcmp(key, "top_levels")) {
        if (value < 1 || value > LRC_TOP_LEVELS_MAX) return lrc_fail(ctx, LRC_ERR_INVALID, "top_levels must be in [1, 8]");
        ctx->opt_top_levels = value;
        return LRC_OK;
    }
    if (!strcmp(key, "scan_chunks")) { if (value < 1) return lrc_fail(ctx, LRC_ERR_INVALID, "scan_chunks must be >= 1"); ctx->opt_scan_chunks = value; return LRC_OK; }
    if (!strcmp(key, "scan_taper")) { if (value < 1) return lrc_fail(ctx, LRC_ERR_INVALID, "scan_taper must be >= 1"); ctx->opt_scan_taper = value; return LRC_OK; }
    if (!strcmp(key, "gather_taper")) { if (value < 1) return lrc_fail(ctx, LRC_ERR_INVALID, "gather_taper must be >= 1"); ctx->opt_gather_taper = value; return LRC_OK; }
    if (!strcmp(key, "gather_ramp")) { if (value < 1) return lrc_fail(ctx, LRC_ERR_INVALID, "gather_ramp must be >= 1"); ctx->opt_gather_ramp = value; return LRC_OK; }
    if (!strcmp(key, "push_tile")) {
        if (value < 2048 || value > TMA_TILE_MAX || (value & 2047)) return lrc_fail(ctx, LRC_ERR_INVALID, "push_tile must be a multiple of 2048 in [2048, 16384]");
        ctx->opt_push_tile = value;
        return LRC_OK;
    }
    if (!strcmp(key, "push_mode")) { if (value < 0 || value > 1) return lrc_fail(ctx, LRC_ERR_INVALID, "push_mode must be 0 (LSU kernel) or 1 (TMA bulk copies)"); ctx->opt_push_mode = value; return LRC_OK; }
    if (!strcmp(key, "push_blocks")) { if (value < 1 || value > 1024) return lrc_fail(ctx, LRC_ERR_INVALID, "push_blocks must be in [1, 1024]"); ctx->opt_push_blocks = value; return LRC_OK; }
    if (!strcmp(key, "gather_chunks")) { if (value < 1) return lrc_fail(ctx, LRC_ERR_INVALID, "gather_chunks must be >= 1"); ctx->opt_gather_chunks = value; return LRC_OK; }
    if (!strcmp(key, "block")) {
        if (value != 0 && value != 32 && value != 64 && value != 128) return lrc_fail(ctx, LRC_ERR_INVALID, "block must be 0 (by call size), 32, 64 or 128");
        ctx->opt_block = value;
        return LRC_OK;
    }
    if (!strcmp(key, "rays_per_thread")) {
        if (value != 1 && value != 2 && value != 4) return lrc_fail(ctx, LRC_ERR_INVALID, "rays_per_thread must be 1, 2 or 4");
        ctx->opt_rays_per_thread = value;
        return LRC_OK;
    }
    if (!strcmp(key, "persistent")) {
        if (value < 0 || value > 2) return lrc_fail(ctx, LRC_ERR_INVALID, "persistent must be 0, 1 or 2 (2 = 48-register build)");
        ctx->opt_persistent = value;
        return LRC_OK;
    }
    if (!strcmp(key, "warp_packet")) { ctx->opt_warp_packet = value != 0; return LRC_OK; }
    if (!strcmp(key, "tune")) {
        if (value != 0 && value != 1 && value != 2 && value != 4 && value != 6 && value != 7 && value != 10 && value != 18 && value != 26)
            return lrc_fail(ctx, LRC_ERR_INVALID, "tune must be 0, 1, 2, 4, 6, 7, 10, 18 or 26");
        ctx->opt_tune = value;
        return LRC_OK;
    }
    if (!strcmp(key, "kernel_timing")) { ctx->opt_kernel_timing = value != 0; ctx->kt_used = 0; return LRC_OK; }
    if (!strcmp(key, "variant")) {
        if (value < 0 || (value > 3 && value != 5 && value != 13 && value != 21 && value != 65)) return lrc_fail(ctx, LRC_ERR_INVALID, "variant must be 0..3, 5, 13, 21 or 65");
        ctx->opt_variant = value;
        return LRC_OK;
    }
    return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_option: unknown key '%s'", key);
}

extern "C" int lrc_cast_rays(lrc_ctx* ctx, const float* rays, int64_t N, float* t_hit, uint32_t* prim_id, void* stream)
{
    int rc = precheck(ctx, "lrc_cast_rays");
    if (rc) return rc;
    if (N < 0 || (N > 0 && (!rays || !t_hit || !prim_id))) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_cast_rays: bad arguments");
    if ((rc = ensure_counters(ctx))) return rc;
    RayGen g;
    memset(&g, 0, sizeof g);
    g.rays = rays; g.N = 1;
    FrameMath fm; fm.max_range = -1.0; fm.cx = fm.cy = fm.cz = 0.0;
    return launch_trace<MODE_RAYS, true>(ctx, g, fm, N, nullptr, nullptr, nullptr, t_hit, prim_id, (cudaStream_t)stream);
}

extern "C" int lrc_cast_rays_bruteforce(lrc_ctx* ctx, const float* rays, int64_t N, float* t_hit, uint32_t* prim_id, void* stream)
{
    int rc = precheck(ctx, "lrc_cast_rays_bruteforce");
    if (rc) return rc;
    if (N < 0 || (N > 0 && (!rays || !t_hit || !prim_id))) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_cast_rays_bruteforce: bad arguments");
    if (N == 0) return LRC_OK;
    k_brute<<<(unsigned)((N + 127) / 128), 128, 0, (cudaStream_t)stream>>>(rays, N, ctx->tris, ctx->T, t_hit, prim_id);
    LRC_CHECK_LAUNCH(ctx, "k_brute");
    return LRC_OK;
}

extern "C" int lrc_rays_intersect(lrc_ctx* ctx, const float* rays, int64_t N, lrc_out* out, void* stream)
{
    int rc = precheck(ctx, "lrc_rays_intersect");
    if (rc) return rc;
    if (N < 0 || (N > 0 && !rays)) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_rays_intersect: bad arguments");
    RayGen g;
    memset(&g, 0, sizeof g);
    g.rays = rays; g.N = (int)(N > 0 ? 1 : 0);
    if (N >= (int64_t)1 << 31) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_rays_intersect: N must be < 2^31 per call");
    // one frame of N rays
    return run_scan<MODE_RAYS>(ctx, g, 1, N, nullptr, -1.0, out, (cudaStream_t)stream);
}

extern "C" int lrc_scan_rays(lrc_ctx* ctx, const float* rays, int64_t N, const double* h_center, double max_range,
                             lrc_out* out, void* stream)
{
    int rc = precheck(ctx, "lrc_scan_rays");
    if (rc) return rc;
    if (N < 0 || (N > 0 && !rays) || !h_center) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_scan_rays: bad arguments");
    if (N >= (int64_t)1 << 31) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_scan_rays: N must be < 2^31 per call");
    RayGen g;
    memset(&g, 0, sizeof g);
    g.rays = rays; g.N = 1;
    return run_scan<MODE_RAYS>(ctx, g, 1, N, h_center, max_range, out, (cudaStream_t)stream);
}

extern "C" int lrc_scan_single_axis(lrc_ctx* ctx, const double* poses, int64_t P, const lrc_single_axis* s,
                                    const lrc_noise* nz, lrc_out* out, void* stream)
{
    int rc = precheck(ctx, "lrc_scan_single_axis");
    if (rc) return rc;
    if (P < 0 || (P > 0 && !poses)) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_scan_single_axis: bad poses");
    RayGen g;
    if ((rc = fill_single(ctx, s, poses, g, (cudaStream_t)stream))) return rc;
    fill_noise(nz, g, false);
    return run_scan<MODE_SINGLE>(ctx, g, P, g.N, nullptr, s->max_range, out, (cudaStream_t)stream);
}

extern "C" int lrc_scan_dual_axis(lrc_ctx* ctx, const double* poses, int64_t P, const lrc_dual_axis* s,
                                  const lrc_noise* nz, lrc_out* out, void* stream)
{
    int rc = precheck(ctx, "lrc_scan_dual_axis");
    if (rc) return rc;
    if (P < 0 || (P > 0 && !poses)) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_scan_dual_axis: bad poses");
    RayGen g;
    if ((rc = fill_dual(ctx, s, poses, g))) return rc;
    fill_noise(nz, g, true);
    return run_scan<MODE_DUAL>(ctx, g, P, g.N, nullptr, s->max_range, out, (cudaStream_t)stream);
}

extern "C" int lrc_gen_rays_single_axis(lrc_ctx* ctx, const double* poses, int64_t P, const lrc_single_axis* s,
                                        float* rays, void* stream)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_gen_rays_single_axis: ctx is NULL");
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (P < 0 || (P > 0 && (!poses || !rays))) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_gen_rays_single_axis: bad arguments");
    RayGen g;
    int rc = fill_single(ctx, s, poses, g, (cudaStream_t)stream);
    if (rc) return rc;
    const int64_t n = P * g.N;
    if (n == 0) return LRC_OK;
    k_gen_rays<MODE_SINGLE><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, n, rays, nullptr);
    LRC_CHECK_LAUNCH(ctx, "k_gen_rays");
    return LRC_OK;
}

extern "C" int lrc_gen_rays_dual_axis(lrc_ctx* ctx, const double* poses, int64_t P, const lrc_dual_axis* s,
                                      const lrc_noise* nz, float* rays, uint8_t* keep, void* stream)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_gen_rays_dual_axis: ctx is NULL");
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (P < 0 || (P > 0 && (!poses || !rays))) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_gen_rays_dual_axis: bad arguments");
    RayGen g;
    int rc = fill_dual(ctx, s, poses, g);
    if (rc) return rc;
    fill_noise(nz, g, true);
    const int64_t n = P * g.N;
    if (n == 0) return LRC_OK;
    k_gen_rays<MODE_DUAL><<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(g, n, rays, keep);
    LRC_CHECK_LAUNCH(ctx, "k_gen_rays");
    return LRC_OK;
}

// ======================================================================================================
// Host-buffer entry points: chunked scan with the device-to-host copies pipelined behind later chunks.
namespace {

int ensure_host_plumbing(lrc_ctx* ctx, size_t n_chunks, int64_t P)
{
    if (!ctx->s_compute) {
        LRC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_compute, cudaStreamNonBlocking));
        LRC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking));
        LRC_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->s_count, cudaStreamNonBlocking));
    }
    if (ctx->n_events < 2 * n_chunks) {
        cudaEvent_t* ev = (cudaEvent_t*)realloc(ctx->events, sizeof(cudaEvent_t) * 2 * n_chunks);
        if (!ev) return lrc_fail(ctx, LRC_ERR_CUDA, "out of host memory for events");
        ctx->events = ev;
        for (size_t i = ctx->n_events; i < 2 * n_chunks; ++i) LRC_CUDA(ctx, cudaEventCreateWithFlags(&ctx->events[i], cudaEventDisableTiming));
        ctx->n_events = 2 * n_chunks;
    }
    const size_t need = (size_t)P + n_chunks + 8;
    if (ctx->h_stage_elems < need) {
        if (ctx->h_stage) LRC_CUDA(ctx, cudaFreeHost(ctx->h_stage));
        ctx->h_stage = nullptr; ctx->h_stage_elems = 0;
        LRC_CUDA(ctx, cudaHostAlloc((void**)&ctx->h_stage, sizeof(int64_t) * need, cudaHostAllocDefault));
        ctx->h_stage_elems = need;
    }
    return LRC_OK;
}

template <int MODE>
int run_scan_host(lrc_ctx* ctx, RayGen g, const double* h_poses, int64_t P, int64_t N, double max_range, lrc_out* h_out,
                  int64_t chunk_poses, int64_t* h_num_points)
{
    if (!h_out || !h_out->xyz || !h_out->frame_offset || !h_num_points)
        return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_scan_*_host: xyz, frame_offset and h_num_points are required");
    if (h_out->capacity < P * N) return lrc_fail(ctx, LRC_ERR_CAPACITY, "lrc_scan_*_host: capacity is smaller than the number of rays");
    *h_num_points = 0;
    h_out->frame_offset[0] = 0;
    if (P == 0) return LRC_OK;
    // Chunk plan.  The PCIe copy is the long pole (3x slower than the kernels), so it should start early and then move
    // few, large pieces (one 307 MB copy runs at 57 GB/s, 8 MB pieces at 52.7): automatic mode (chunk_poses <= 0) starts
    // with ~0.5M rays and doubles the chunk every time; an explicit chunk_poses gives uniform chunks.
    std::vector<int64_t> chunk_start;
    {
        const bool grow = chunk_poses <= 0;
        int64_t cp = grow ? (int64_t)(1 << 19) / N : chunk_poses;
        if (cp < 1) cp = 1;
        for (int64_t f = 0; f < P;) {
            chunk_start.push_back(f);
            f += cp;
            if (grow) cp *= 2;
        }
        chunk_start.push_back(P);
    }
    const size_t n_chunks = chunk_start.size() - 1;
    int rc = ensure_host_plumbing(ctx, n_chunks, P);
    if (rc) return rc;
    // device image of the whole trajectory's outputs: [poses | xyz | incident | prim | label | ray | frame offsets]
    const size_t cap = (size_t)(P * N);
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_pose = carve(sizeof(double) * 16 * (size_t)P);
    const size_t o_xyz = carve(sizeof(float) * 3 * cap);
    const size_t o_inc = h_out->incident_deg ? carve(sizeof(double) * cap) : 0;
    const size_t o_prim = h_out->prim_id ? carve(sizeof(uint32_t) * cap) : 0;
    const size_t o_lab = h_out->label ? carve(sizeof(uint32_t) * cap) : 0;
    const size_t o_ray = h_out->ray_idx ? carve(sizeof(uint32_t) * cap) : 0;
    const size_t o_off = carve(sizeof(int64_t) * ((size_t)P + n_chunks));
    if ((rc = lrc_grow(ctx, &ctx->host_dev, &ctx->host_dev_bytes, off))) return rc;
    char* base = (char*)ctx->host_dev;
    double* d_poses = (double*)(base + o_pose);
    LRC_CUDA(ctx, cudaMemcpyAsync(d_poses, h_poses, sizeof(double) * 16 * (size_t)P, cudaMemcpyHostToDevice, ctx->s_compute));
    g.poses = d_poses;
    const uint64_t pose_base = g.pose_index_base;
    // One small frame (the reference's per-waypoint call pattern): a count round trip would double the latency, so the
    // records are copied at capacity -- at most 6 MB -- right behind the kernels and one synchronisation ends the call.
    if (n_chunks == 1 && P * N <= ((int64_t)1 << 18)) {
        lrc_out d;
        d.xyz = (float*)(base + o_xyz);
        d.incident_deg = h_out->incident_deg ? (double*)(base + o_inc) : nullptr;
        d.prim_id = h_out->prim_id ? (uint32_t*)(base + o_prim) : nullptr;
        d.label = h_out->label ? (uint32_t*)(base + o_lab) : nullptr;
        d.ray_idx = h_out->ray_idx ? (uint32_t*)(base + o_ray) : nullptr;
        d.frame_offset = (int64_t*)(base + o_off);
        d.capacity = P * N;
        cudaStream_t st = ctx->s_compute;
        if ((rc = run_scan<MODE>(ctx, g, P, N, nullptr, max_range, &d, st))) return rc;
        LRC_CUDA(ctx, cudaMemcpyAsync(ctx->h_stage, d.frame_offset, sizeof(int64_t) * (size_t)(P + 1), cudaMemcpyDeviceToHost, st));
        LRC_CUDA(ctx, cudaMemcpyAsync(h_out->xyz, d.xyz, sizeof(float) * 3 * cap, cudaMemcpyDeviceToHost, st));
        if (h_out->incident_deg) LRC_CUDA(ctx, cudaMemcpyAsync(h_out->incident_deg, d.incident_deg, sizeof(double) * cap, cudaMemcpyDeviceToHost, st));
        if (h_out->label) LRC_CUDA(ctx, cudaMemcpyAsync(h_out->label, d.label, sizeof(uint32_t) * cap, cudaMemcpyDeviceToHost, st));
        if (h_out->prim_id) LRC_CUDA(ctx, cudaMemcpyAsync(h_out->prim_id, d.prim_id, sizeof(uint32_t) * cap, cudaMemcpyDeviceToHost, st));
        if (h_out->ray_idx) LRC_CUDA(ctx, cudaMemcpyAsync(h_out->ray_idx, d.ray_idx, sizeof(uint32_t) * cap, cudaMemcpyDeviceToHost, st));
        LRC_CUDA(ctx, cudaStreamSynchronize(st));
        for (int64_t k = 0; k <= P; ++k) h_out->frame_offset[k] = ctx->h_stage[k];
        *h_num_points = ctx->h_stage[P];
        return LRC_OK;
    }
    // 1) enqueue every chunk's kernels; a tiny copy of each chunk's frame offsets follows on its own stream
    for (size_t c = 0; c < n_chunks; ++c) {
        const int64_t f0 = chunk_start[c];
        const int64_t nf = chunk_start[c + 1] - f0;
        lrc_out d;
        d.xyz = (float*)(base + o_xyz) + 3 * (size_t)(f0 * N);
        d.incident_deg = h_out->incident_deg ? (double*)(base + o_inc) + (size_t)(f0 * N) : nullptr;
        d.prim_id = h_out->prim_id ? (uint32_t*)(base + o_prim) + (size_t)(f0 * N) : nullptr;
        d.label = h_out->label ? (uint32_t*)(base + o_lab) + (size_t)(f0 * N) : nullptr;
        d.ray_idx = h_out->ray_idx ? (uint32_t*)(base + o_ray) + (size_t)(f0 * N) : nullptr;
        d.frame_offset = (int64_t*)(base + o_off) + f0 + (int64_t)c;     // nf + 1 entries per chunk
        d.capacity = nf * N;
        RayGen gc = g;
        gc.poses = d_poses + 16 * f0;
        gc.pose_index_base = pose_base + (uint64_t)f0;
        if ((rc = run_scan<MODE>(ctx, gc, nf, N, nullptr, max_range, &d, ctx->s_compute))) return rc;
        LRC_CUDA(ctx, cudaEventRecord(ctx->events[2 * c], ctx->s_compute));
        LRC_CUDA(ctx, cudaStreamWaitEvent(ctx->s_count, ctx->events[2 * c], 0));
        LRC_CUDA(ctx, cudaMemcpyAsync(ctx->h_stage + f0 + (int64_t)c, d.frame_offset, sizeof(int64_t) * (size_t)(nf + 1),
                                      cudaMemcpyDeviceToHost, ctx->s_count));
        LRC_CUDA(ctx, cudaEventRecord(ctx->events[2 * c + 1], ctx->s_count));
    }
    // 2) as each chunk's counts arrive, enqueue exactly-sized record copies on the copy stream
    int64_t total = 0;
    for (size_t c = 0; c < n_chunks; ++c) {
        const int64_t f0 = chunk_start[c];
        const int64_t nf = chunk_start[c + 1] - f0;
        LRC_CUDA(ctx, cudaEventSynchronize(ctx->events[2 * c + 1]));
        const int64_t* st = ctx->h_stage + f0 + (int64_t)c;
        const int64_t m = st[nf];
        for (int64_t k = 0; k <= nf; ++k) h_out->frame_offset[f0 + k] = total + st[k];
        if (m > 0) {
            const size_t src = (size_t)(f0 * N);
            LRC_CUDA(ctx, cudaStreamWaitEvent(ctx->s_copy, ctx->events[2 * c], 0));
            LRC_CUDA(ctx, cudaMemcpyAsync(h_out->xyz + 3 * total, (float*)(base + o_xyz) + 3 * src, sizeof(float) * 3 * (size_t)m, cudaMemcpyDeviceToHost, ctx->s_copy));
            if (h_out->incident_deg)
                LRC_CUDA(ctx, cudaMemcpyAsync(h_out->incident_deg + total, (double*)(base + o_inc) + src, sizeof(double) * (size_t)m, cudaMemcpyDeviceToHost, ctx->s_copy));
            if (h_out->label)
                LRC_CUDA(ctx, cudaMemcpyAsync(h_out->label + total, (uint32_t*)(base + o_lab) + src, sizeof(uint32_t) * (size_t)m, cudaMemcpyDeviceToHost, ctx->s_copy));
            if (h_out->prim_id)
                LRC_CUDA(ctx, cudaMemcpyAsync(h_out->prim_id + total, (uint32_t*)(base + o_prim) + src, sizeof(uint32_t) * (size_t)m, cudaMemcpyDeviceToHost, ctx->s_copy));
            if (h_out->ray_idx)
                LRC_CUDA(ctx, cudaMemcpyAsync(h_out->ray_idx + total, (uint32_t*)(base + o_ray) + src, sizeof(uint32_t) * (size_t)m, cudaMemcpyDeviceToHost, ctx->s_copy));
        }
        total += m;
    }
    LRC_CUDA(ctx, cudaStreamSynchronize(ctx->s_copy));
    LRC_CUDA(ctx, cudaStreamSynchronize(ctx->s_compute));
    *h_num_points = total;
    return LRC_OK;
}

}  // namespace

extern "C" int lrc_scan_single_axis_host(lrc_ctx* ctx, const double* h_poses, int64_t P, const lrc_single_axis* s,
                                         const lrc_noise* nz, lrc_out* h_out, int64_t chunk_poses, int64_t* h_num_points)
{
    int rc = precheck(ctx, "lrc_scan_single_axis_host");
    if (rc) return rc;
    if (P < 0 || (P > 0 && !h_poses)) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_scan_single_axis_host: bad poses");
    if ((rc = ensure_host_plumbing(ctx, 1, P > 0 ? P : 1))) return rc;
    RayGen g;
    if ((rc = fill_single(ctx, s, nullptr, g, ctx->s_compute))) return rc;
    fill_noise(nz, g, false);
    return run_scan_host<MODE_SINGLE>(ctx, g, h_poses, P, g.N, s->max_range, h_out, chunk_poses, h_num_points);
}

extern "C" int lrc_scan_dual_axis_host(lrc_ctx* ctx, const double* h_poses, int64_t P, const lrc_dual_axis* s,
                                       const lrc_noise* nz, lrc_out* h_out, int64_t chunk_poses, int64_t* h_num_points)
{
    int rc = precheck(ctx, "lrc_scan_dual_axis_host");
    if (rc) return rc;
    if (P < 0 || (P > 0 && !h_poses)) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_scan_dual_axis_host: bad poses");
    if ((rc = ensure_host_plumbing(ctx, 1, P > 0 ? P : 1))) return rc;
    RayGen g;
    if ((rc = fill_dual(ctx, s, nullptr, g))) return rc;
    fill_noise(nz, g, true);
    return run_scan_host<MODE_DUAL>(ctx, g, h_poses, P, g.N, s->max_range, h_out, chunk_poses, h_num_points);
}

extern "C" int lrc_set_mesh_host(lrc_ctx* ctx, const float* h_verts, int64_t V, const int32_t* h_tris, int64_t T,
                                 const uint32_t* h_tri_label)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_set_mesh_host: ctx is NULL");
    if (V < 0 || T < 0 || (T > 0 && (!h_verts || !h_tris))) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_mesh_host: bad arguments");
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = ensure_host_plumbing(ctx, 1, 1);
    if (rc) return rc;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_v = carve(sizeof(float) * 3 * (size_t)V), o_t = carve(sizeof(int32_t) * 3 * (size_t)T);
    const size_t o_l = h_tri_label ? carve(sizeof(uint32_t) * (size_t)T) : 0;
    if ((rc = lrc_grow(ctx, &ctx->mesh_dev, &ctx->mesh_dev_bytes, off))) return rc;
    char* base = (char*)ctx->mesh_dev;
    cudaStream_t st = ctx->s_compute;
    if (V > 0) LRC_CUDA(ctx, cudaMemcpyAsync(base + o_v, h_verts, sizeof(float) * 3 * (size_t)V, cudaMemcpyHostToDevice, st));
    if (T > 0) LRC_CUDA(ctx, cudaMemcpyAsync(base + o_t, h_tris, sizeof(int32_t) * 3 * (size_t)T, cudaMemcpyHostToDevice, st));
    if (h_tri_label && T > 0) LRC_CUDA(ctx, cudaMemcpyAsync(base + o_l, h_tri_label, sizeof(uint32_t) * (size_t)T, cudaMemcpyHostToDevice, st));
    return lrc_set_mesh(ctx, (const float*)(base + o_v), V, (const int32_t*)(base + o_t), T,
                        h_tri_label ? (const uint32_t*)(base + o_l) : nullptr, st);
}

// ======================================================================================================
// Multi-GPU exchange over NVLink peer memory: buffers that other processes map with CUDA IPC, and the gather
// targets the compaction kernel stores into.
extern "C" int lrc_peer_buffer_create(lrc_ctx* ctx, int64_t bytes, void** d_ptr, lrc_ipc_handle* h_handle)
{
    if (!ctx || !d_ptr || !h_handle || bytes <= 0) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_peer_buffer_create: bad arguments");
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(lrc_ipc_handle), "handle size");
    void* p = nullptr;
    LRC_CUDA(ctx, cudaMalloc(&p, (size_t)bytes));
    LRC_CUDA(ctx, cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return lrc_fail(ctx, LRC_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    memset(h_handle, 0, sizeof *h_handle);
    memcpy(h_handle, &h, sizeof h);
    *d_ptr = p;
    return LRC_OK;
}

extern "C" int lrc_peer_buffer_open(lrc_ctx* ctx, const lrc_ipc_handle* h_handle, void** d_ptr)
{
    if (!ctx || !d_ptr || !h_handle) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_peer_buffer_open: bad arguments");
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, h_handle, sizeof h);
    LRC_CUDA(ctx, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return LRC_OK;
}

extern "C" int lrc_peer_buffer_close(lrc_ctx* ctx, void* d_ptr)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_peer_buffer_close: ctx is NULL");
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (d_ptr) LRC_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
    return LRC_OK;
}

extern "C" int lrc_peer_buffer_destroy(lrc_ctx* ctx, void* d_ptr)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_peer_buffer_destroy: ctx is NULL");
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (d_ptr) LRC_CUDA(ctx, cudaFree(d_ptr));
    return LRC_OK;
}

extern "C" int lrc_set_gather_wire(lrc_ctx* ctx, const lrc_gather_wire* w)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_set_gather_wire: ctx is NULL");
    memset(&ctx->wire, 0, sizeof ctx->wire);
    ctx->wire_scan = 0;
    if (!w || !w->enabled) return LRC_OK;
    if (ctx->gather.n <= 0) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_gather_wire: call lrc_set_gather first");
    if (w->self < 0 || w->self >= ctx->gather.n || !w->all_poses) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_gather_wire: bad self / all_poses");
    for (int k = 0; k < ctx->gather.n; ++k) {
        if (!w->t[k] || !w->ray_idx[k] || !w->ready[k]) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_gather_wire: every target needs t, ray_idx and ready");
        if (w->rank_frames[k] < 0 || w->rank_pose0[k] < 0 || w->rank_point_base[k] < 0 || w->rank_frame_base[k] < 0)
            return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_gather_wire: negative rank table entry");
        if (w->rank_frames[k] >= ((int64_t)1 << 31)) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_gather_wire: too many frames");
    }
    if (w->rank_point_base[w->self] != ctx->gather.point_base || w->rank_frame_base[w->self] != ctx->gather.frame_base)
        return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_gather_wire: rank tables disagree with lrc_set_gather for this rank");
    ctx->wire = *w;
    return LRC_OK;
}

extern "C" int lrc_set_gather(lrc_ctx* ctx, const lrc_gather* h_targets)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_set_gather: ctx is NULL");
    memset(&ctx->wire, 0, sizeof ctx->wire);          // the compact wire format belongs to one set of targets
    ctx->wire_scan = 0;
    if (!h_targets || h_targets->n_targets == 0) { memset(&ctx->gather, 0, sizeof ctx->gather); return LRC_OK; }
    if (h_targets->n_targets < 0 || h_targets->n_targets > LRC_MAX_GATHER) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_gather: n_targets out of range");
    for (int k = 0; k < h_targets->n_targets; ++k)
        if (!h_targets->xyz[k] || !h_targets->label[k] || !h_targets->frame_offset[k])
            return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_gather: every target needs xyz, label and frame_offset");
    if (h_targets->point_base < 0 || h_targets->frame_base < 0 || h_targets->capacity < 0 || h_targets->frame_capacity < 0) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_set_gather: negative base");
    ctx->gather.n = h_targets->n_targets;
    for (int k = 0; k < h_targets->n_targets; ++k) {
        ctx->gather.xyz[k] = h_targets->xyz[k]; ctx->gather.label[k] = h_targets->label[k]; ctx->gather.frame_offset[k] = h_targets->frame_offset[k];
    }
    ctx->gather.point_base = h_targets->point_base; ctx->gather.frame_base = h_targets->frame_base; ctx->gather.capacity = h_targets->capacity;
    ctx->gather.frame_capacity = h_targets->frame_capacity;
    return LRC_OK;
}
