// post.cu -- the two steps that directly follow the ray-cast call in the reference's frame loop and export
// (SURVEY.md section 8f-4 and 8f-2), kept on the GPU so that the compacted cloud does not have to visit the host
// before it is summarised and serialised:
//
//   lrc_frame_statistics   per-frame ScanQuality sums              (reference s3dis_simulator.py:276-284)
//   lrc_pack_ply_records   19-byte labelled-PLY vertex records     (reference containers/s3dis_sim_scene.py:614-641)
//
// Both are streaming, HBM-bound passes over the compacted SoA output of a scan (lrc_out).
#include "common.cuh"

namespace {

// ---- f-4: ScanQuality ---------------------------------------------------------------------------------
// The reference computes, per frame (s3dis_simulator.py:276-284):
//     num_points = len(points)
//     incident_angle_mean/std = np.mean / np.std of the float64 incident angles         (population std)
//     range_mean/std          = np.mean / np.std of np.linalg.norm(points, axis=1)      (float32 norms, taken from
//                               the WORLD ORIGIN, not from the sensor -- reproduced as is)
// Here: each float32 norm is evaluated exactly as numpy does (sqrt((x*x + y*y) + z*z) in float32), sums are carried
// in float64 in a fixed order (thread-strided partials -> shuffle tree -> warp 0 -> S partial records per frame ->
// one thread per frame), so the result is deterministic and independent of the number of GPUs.
constexpr int FS_THREADS = 256;
constexpr int FS_SPLITS = 8;

struct Partial { double n, sa, saa, sr, srr; };

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(FS_THREADS)
k_frame_stats_partial(const float* __restrict__ xyz, const double* __restrict__ inc, const int64_t* __restrict__ frame_offset,
                      int64_t P, Partial* __restrict__ part)
{
    const int64_t frame = blockIdx.x / FS_SPLITS;
    const int split = blockIdx.x % FS_SPLITS;
    const int64_t a = frame_offset[frame], b = frame_offset[frame + 1];
    const int64_t m = b - a;
    const int64_t per = (m + FS_SPLITS - 1) / FS_SPLITS;
    const int64_t s0 = a + per * split;
    const int64_t s1 = s0 + per < b ? s0 + per : b;
    double sa = 0.0, saa = 0.0, sr = 0.0, srr = 0.0;
    for (int64_t i = s0 + threadIdx.x; i < s1; i += FS_THREADS) {
        const float x = __ldcs(xyz + 3 * i), y = __ldcs(xyz + 3 * i + 1), z = __ldcs(xyz + 3 * i + 2);
        const float r = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
        const double rd = (double)r;
        sr += rd;
        srr += rd * rd;
        if (inc) {
            const double v = __ldcs(inc + i);
            sa += v;
            saa += v * v;
        }
    }
    __shared__ double sh[4][FS_THREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    sa = warp_sum(sa); saa = warp_sum(saa); sr = warp_sum(sr); srr = warp_sum(srr);
    if (lane == 0) { sh[0][w] = sa; sh[1][w] = saa; sh[2][w] = sr; sh[3][w] = srr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        Partial p = {(double)(s1 > s0 ? s1 - s0 : 0), 0.0, 0.0, 0.0, 0.0};
        for (int k = 0; k < FS_THREADS / 32; ++k) { p.sa += sh[0][k]; p.saa += sh[1][k]; p.sr += sh[2][k]; p.srr += sh[3][k]; }
        part[blockIdx.x] = p;
    }
}

__global__ void k_frame_stats_final(const Partial* __restrict__ part, const int64_t* __restrict__ frame_offset, int64_t P,
                                    lrc_frame_stats* __restrict__ out)
{
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= P) return;
    double sa = 0.0, saa = 0.0, sr = 0.0, srr = 0.0;
    for (int k = 0; k < FS_SPLITS; ++k) {
        const Partial p = part[f * FS_SPLITS + k];
        sa += p.sa; saa += p.saa; sr += p.sr; srr += p.srr;
    }
    const int64_t m = frame_offset[f + 1] - frame_offset[f];
    lrc_frame_stats s;
    s.num_points = m;
    if (m > 0) {
        const double n = (double)m;
        s.incident_mean = sa / n;
        s.incident_std = sqrt(fmax(saa / n - s.incident_mean * s.incident_mean, 0.0));
        s.range_mean = sr / n;
        s.range_std = sqrt(fmax(srr / n - s.range_mean * s.range_mean, 0.0));
    } else {   // reference: "... if len(...) > 0 else 0"
        s.incident_mean = s.incident_std = s.range_mean = s.range_std = 0.0;
    }
    out[f] = s;
}

// ---- f-2: labelled PLY records ------------------------------------------------------------------------
// Record layout of the reference writer (s3dis_sim_scene.py:634-641), little endian, 19 bytes, no padding:
//     float32 x, y, z | uint8 red, green, blue | uint16 sem | uint16 ins
// A block packs PLY_POINTS records into shared memory and streams them out as 16-byte vectors: the block's byte range
// starts at a multiple of 19 * PLY_POINTS = 4864 = 304 * 16, so every vector store is aligned and coalesced.
constexpr int PLY_POINTS = 256;
constexpr int PLY_REC = 19;

__global__ void __launch_bounds__(PLY_POINTS)
k_pack_ply(const float* __restrict__ xyz, const uint32_t* __restrict__ label, const uint32_t* __restrict__ prim_id,
           const uint32_t* __restrict__ tri_rgb, uint32_t default_rgb, int64_t M, uint8_t* __restrict__ out)
{
    __shared__ __align__(16) uint8_t sh[PLY_POINTS * PLY_REC];
    const int64_t i = (int64_t)blockIdx.x * PLY_POINTS + threadIdx.x;
    if (i < M) {
        const float x = __ldcs(xyz + 3 * i), y = __ldcs(xyz + 3 * i + 1), z = __ldcs(xyz + 3 * i + 2);
        const uint32_t lab = label ? __ldcs(label + i) : 0u;
        uint32_t rgb = default_rgb;
        if (tri_rgb && prim_id) rgb = __ldg(tri_rgb + __ldcs(prim_id + i));
        uint8_t* r = sh + threadIdx.x * PLY_REC;
        const uint32_t w[3] = {__float_as_uint(x), __float_as_uint(y), __float_as_uint(z)};
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            r[4 * k + 0] = (uint8_t)(w[k]); r[4 * k + 1] = (uint8_t)(w[k] >> 8);
            r[4 * k + 2] = (uint8_t)(w[k] >> 16); r[4 * k + 3] = (uint8_t)(w[k] >> 24);
        }
        r[12] = (uint8_t)(rgb); r[13] = (uint8_t)(rgb >> 8); r[14] = (uint8_t)(rgb >> 16);     // red, green, blue
        r[15] = (uint8_t)(lab); r[16] = (uint8_t)(lab >> 8);                                     // sem  (low 16 bits)
        r[17] = (uint8_t)(lab >> 16); r[18] = (uint8_t)(lab >> 24);                              // ins  (high 16 bits)
    }
    __syncthreads();
    const int64_t byte0 = (int64_t)blockIdx.x * (PLY_POINTS * PLY_REC);
    const int64_t total = M * PLY_REC;
    const int64_t nbytes = (total - byte0) < (int64_t)(PLY_POINTS * PLY_REC) ? (total - byte0) : (int64_t)(PLY_POINTS * PLY_REC);
    const int nvec = (int)(nbytes / 16);
    const uint4* src = reinterpret_cast<const uint4*>(sh);
    uint4* dst = reinterpret_cast<uint4*>(out + byte0);
    for (int v = threadIdx.x; v < nvec; v += PLY_POINTS) __stcs(dst + v, src[v]);
    for (int b = nvec * 16 + threadIdx.x; b < nbytes; b += PLY_POINTS) out[byte0 + b] = sh[b];   // tail of the last block
}

}  // namespace

// ---- incident angles from (point, pose) ----------------------------------------------------------------------------------
// The reference's "incident angle" is a pure function of the float32 hit point and the frame's sensor position
// (raycast_engine_cpu.py:95-107): degrees(arccos(|dz / dist|)) with dist = ||p - c|| in float64.  The multi-GPU exchange
// therefore moves xyz + label only (16 B per point instead of 24) and every rank recomputes the angles of the gathered
// cloud on arrival -- same operations in the same order as the scan's own epilogue (scan.cu frame_epilogue), so the bits
// are identical.  One thread per point; the frame of a point is found by bisection of frame_offset.
namespace {
__global__ void k_incident_from_points(const float* __restrict__ xyz, const int64_t* __restrict__ frame_offset, int64_t P,
                                       const double* __restrict__ poses, int64_t M, double* __restrict__ out)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    // offsets may carry a base (a rank's region of the gather buffer starts at rank * capacity): point i of xyz is
    // absolute point frame_offset[0] + i, and the cloud ends at frame_offset[P] whatever upper bound M the caller gave
    const int64_t ia = i + frame_offset[0];
    if (ia >= frame_offset[P]) return;
    int64_t lo = 0, hi = P;                     // largest f with frame_offset[f] <= ia
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (frame_offset[mid] <= ia) lo = mid; else hi = mid;
    }
    const double* Mx = poses + 16 * lo;
    const double cx = Mx[3], cy = Mx[7], cz = Mx[11];
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    const double ddx = __dsub_rn((double)x, cx), ddy = __dsub_rn((double)y, cy), ddz = __dsub_rn((double)z, cz);
    const double dist = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(ddx, ddx), __dmul_rn(ddy, ddy)), __dmul_rn(ddz, ddz)));
    out[i] = __dmul_rn(acos(fabs(__ddiv_rn(ddz, dist))), 180.0 / 3.141592653589793);
}
}  // namespace

extern "C" int lrc_incident_angles(lrc_ctx* ctx, const float* xyz, const int64_t* frame_offset, int64_t P, const double* poses,
                                   int64_t M, double* incident_deg, void* stream)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_incident_angles: ctx is NULL");
    if (M < 0 || P < 0 || (M > 0 && (!xyz || !frame_offset || !poses || !incident_deg || P == 0)))
        return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_incident_angles: bad arguments");
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (M == 0) return LRC_OK;
    k_incident_from_points<<<(unsigned)((M + 255) / 256), 256, 0, (cudaStream_t)stream>>>(xyz, frame_offset, P, poses, M, incident_deg);
    LRC_CHECK_LAUNCH(ctx, "k_incident_from_points");
    return LRC_OK;
}

extern "C" int lrc_frame_statistics(lrc_ctx* ctx, const float* xyz, const double* incident_deg, const int64_t* frame_offset,
                                    int64_t P, lrc_frame_stats* out, void* stream_)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_frame_statistics: ctx is NULL");
    if (P < 0 || (P > 0 && (!xyz || !frame_offset || !out))) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_frame_statistics: bad arguments");
    if (P == 0) return LRC_OK;
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = lrc_grow(ctx, &ctx->post_scratch, &ctx->post_scratch_bytes, sizeof(Partial) * (size_t)P * FS_SPLITS);
    if (rc) return rc;
    Partial* part = (Partial*)ctx->post_scratch;
    k_frame_stats_partial<<<(unsigned)(P * FS_SPLITS), FS_THREADS, 0, stream>>>(xyz, incident_deg, frame_offset, P, part);
    LRC_CHECK_LAUNCH(ctx, "k_frame_stats_partial");
    k_frame_stats_final<<<(unsigned)((P + 127) / 128), 128, 0, stream>>>(part, frame_offset, P, out);
    LRC_CHECK_LAUNCH(ctx, "k_frame_stats_final");
    return LRC_OK;
}

extern "C" int lrc_pack_ply_records(lrc_ctx* ctx, const float* xyz, const uint32_t* label, const uint32_t* prim_id,
                                    const uint32_t* tri_rgb, uint32_t default_rgb, int64_t M, uint8_t* out, void* stream_)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_pack_ply_records: ctx is NULL");
    if (M < 0 || (M > 0 && (!xyz || !out))) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_pack_ply_records: bad arguments");
    if (tri_rgb && !prim_id) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_pack_ply_records: tri_rgb needs prim_id");
    if (((uintptr_t)out & 15u) != 0) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_pack_ply_records: out must be 16-byte aligned");
    if (M == 0) return LRC_OK;
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    k_pack_ply<<<(unsigned)((M + PLY_POINTS - 1) / PLY_POINTS), PLY_POINTS, 0, (cudaStream_t)stream_>>>(
        xyz, label, prim_id, tri_rgb, default_rgb, M, out);
    LRC_CHECK_LAUNCH(ctx, "k_pack_ply");
    return LRC_OK;
}
