// nn.cu -- exact 1-nearest-neighbour label / colour transfer: the labelling step the
// reference applies to every frame's hit points before export,
//     NearestNeighbors(n_neighbors=1, algorithm='ball_tree').fit(s3dis_points).kneighbors(points)
//                                               (reference containers/s3dis_sim_scene.py:413-424 and :505-541)
// followed by colours / labels / instances [indices].  scikit-learn's ball tree is third party; its result is defined
// by the metric alone: argmin over the annotated points of the float64 reduced distance
//     rdist = ((qx - px)^2 + (qy - py)^2) + (qz - pz)^2         (sequential sum over the three axes, no FMA)
// with the query promoted from float32.  The GPU evaluates exactly that expression, so indices agree bit for bit except
// where two annotated points are at EXACTLY the same distance (the ball tree then returns whichever its traversal met
// first; this kernel returns the smaller index).
//
// Structure: annotated points are binned once on a uniform 3-D grid (count -> scan -> scatter, like plan.cu's vertex
// index); a query walks cubic shells of cells around its own cell and stops as soon as the best distance found is not
// larger than the distance to the unexplored region.  HBM/L2-bound gather; one thread per query point.
#include "common.cuh"
#include "scan_util.cuh"

namespace {

struct NnMeta { unsigned long long lo[3], hi[3]; int n_finite; };

struct NnGrid { double o[3]; double cell; int nb[3]; };

__device__ __forceinline__ unsigned long long nn_d2ord(double d)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
static inline double nn_ord2d(unsigned long long u)
{
    unsigned long long b = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    double d;
    memcpy(&d, &b, 8);
    return d;
}
__device__ __forceinline__ bool nn_finite(double x, double y, double z) { return (x - x == 0.0) && (y - y == 0.0) && (z - z == 0.0); }

__global__ void k_nn_meta_init(NnMeta* m)
{
    for (int k = 0; k < 3; ++k) { m->lo[k] = ~0ull; m->hi[k] = 0ull; }
    m->n_finite = 0;
}

__global__ void k_nn_bounds(const double* __restrict__ p, int64_t n, NnMeta* m)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double c[3] = {0, 0, 0};
    bool ok = false;
    if (i < n) { c[0] = p[3 * i]; c[1] = p[3 * i + 1]; c[2] = p[3 * i + 2]; ok = nn_finite(c[0], c[1], c[2]); }
    const unsigned any = __ballot_sync(0xffffffffu, ok);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        unsigned long long l = ok ? nn_d2ord(c[k]) : ~0ull, h = ok ? nn_d2ord(c[k]) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l = min(l, __shfl_xor_sync(0xffffffffu, l, o));
            h = max(h, __shfl_xor_sync(0xffffffffu, h, o));
        }
        if (lane_id() == 0 && any) { atomicMin(&m->lo[k], l); atomicMax(&m->hi[k], h); }
    }
    if (lane_id() == 0 && any) atomicAdd(&m->n_finite, __popc(any));
}

__device__ __forceinline__ int nn_bin(double x, double o, double cell, int nb)
{
    const double f = floor((x - o) / cell);
    return f < 0.0 ? 0 : (f >= (double)nb ? nb - 1 : (int)f);
}
__device__ __forceinline__ int64_t nn_cell(const NnGrid& g, int bx, int by, int bz) { return ((int64_t)bz * g.nb[1] + by) * g.nb[0] + bx; }

__global__ void k_nn_count(const double* __restrict__ p, int64_t n, NnGrid g, int* __restrict__ count)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    if (!nn_finite(x, y, z)) return;          // a NaN point can never be the nearest neighbour
    atomicAdd(&count[nn_cell(g, nn_bin(x, g.o[0], g.cell, g.nb[0]), nn_bin(y, g.o[1], g.cell, g.nb[1]), nn_bin(z, g.o[2], g.cell, g.nb[2]))], 1);
}

__global__ void k_nn_scatter(const double* __restrict__ p, int64_t n, NnGrid g, int* __restrict__ cursor, double* __restrict__ sorted,
                             int32_t* __restrict__ sorted_idx)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    if (!nn_finite(x, y, z)) return;
    const int pos = atomicAdd(&cursor[nn_cell(g, nn_bin(x, g.o[0], g.cell, g.nb[0]), nn_bin(y, g.o[1], g.cell, g.nb[1]), nn_bin(z, g.o[2], g.cell, g.nb[2]))], 1);
    sorted[3 * (int64_t)pos] = x; sorted[3 * (int64_t)pos + 1] = y; sorted[3 * (int64_t)pos + 2] = z;
    sorted_idx[pos] = (int32_t)i;             // order inside a cell depends on the atomics; ties are broken by this index
}

__global__ void __launch_bounds__(128)
k_nn_query(const float* __restrict__ q, int64_t M, NnGrid g, const int* __restrict__ start, const double* __restrict__ sorted,
           const int32_t* __restrict__ sorted_idx, int32_t* __restrict__ out_idx, double* __restrict__ out_dist,
           const uint32_t* __restrict__ ref_a, uint32_t* __restrict__ out_a, const uint32_t* __restrict__ ref_b, uint32_t* __restrict__ out_b)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    const double qx = (double)__ldg(q + 3 * i), qy = (double)__ldg(q + 3 * i + 1), qz = (double)__ldg(q + 3 * i + 2);
    double best = INFINITY;
    int32_t best_idx = -1;
    if (nn_finite(qx, qy, qz) && g.nb[0] > 0) {
        const int cx = nn_bin(qx, g.o[0], g.cell, g.nb[0]), cy = nn_bin(qy, g.o[1], g.cell, g.nb[1]), cz = nn_bin(qz, g.o[2], g.cell, g.nb[2]);
        const int rmax = max(max(max(cx, g.nb[0] - 1 - cx), max(cy, g.nb[1] - 1 - cy)), max(cz, g.nb[2] - 1 - cz));
        for (int r = 0; r <= rmax; ++r) {
            // shell of Chebyshev radius r around (cx, cy, cz), clipped to the grid; rows along x are contiguous in memory
            const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nb[2] - 1);
            const int y0 = max(cy - r, 0), y1 = min(cy + r, g.nb[1] - 1);
            for (int bz = z0; bz <= z1; ++bz)
                for (int by = y0; by <= y1; ++by) {
                    const bool face = (bz == cz - r) || (bz == cz + r) || (by == cy - r) || (by == cy + r);
                    // on a z/y face of the shell the whole x-row belongs to it; otherwise only its two end cells
                    for (int part = 0; part < (face ? 1 : 2); ++part) {
                        int xa, xb;
                        if (face) { xa = max(cx - r, 0); xb = min(cx + r, g.nb[0] - 1); }
                        else { xa = xb = part == 0 ? cx - r : cx + r; if (xa < 0 || xa >= g.nb[0] || (part == 1 && r == 0)) continue; }
                        const int s = start[nn_cell(g, xa, by, bz)], e = start[nn_cell(g, xb, by, bz) + 1];
                        for (int k = s; k < e; ++k) {
                            const double dx = __dsub_rn(qx, sorted[3 * (int64_t)k]), dy = __dsub_rn(qy, sorted[3 * (int64_t)k + 1]),
                                         dz = __dsub_rn(qz, sorted[3 * (int64_t)k + 2]);
                            const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                            const int32_t id = sorted_idx[k];
                            if (d < best || (d == best && id < best_idx)) { best = d; best_idx = id; }
                        }
                    }
                }
            if (best_idx >= 0) {
                // every unexplored point lies outside the cube of cells [c - r, c + r]: at least `gap` away (a query
                // outside the grid was clamped into it, which only shrinks the bound, i.e. searches further)
                double gap = INFINITY;
                const double qq[3] = {qx, qy, qz};
                const int cc[3] = {cx, cy, cz};
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    if (cc[a] - r > 0) gap = fmin(gap, qq[a] - (g.o[a] + (double)(cc[a] - r) * g.cell));
                    if (cc[a] + r < g.nb[a] - 1) gap = fmin(gap, (g.o[a] + (double)(cc[a] + r + 1) * g.cell) - qq[a]);
                }
                gap = fmax(gap, 0.0) * (1.0 - 1e-9);      // bin edges are computed in floating point: stay conservative
                if (best <= gap * gap) break;
            }
        }
    }
    out_idx[i] = best_idx;
    if (out_dist) out_dist[i] = best_idx >= 0 ? __dsqrt_rn(best) : INFINITY;
    if (out_a) out_a[i] = (ref_a && best_idx >= 0) ? __ldg(ref_a + best_idx) : 0u;
    if (out_b) out_b[i] = (ref_b && best_idx >= 0) ? __ldg(ref_b + best_idx) : 0u;
}

}  // namespace

extern "C" int lrc_nn_index_build(lrc_ctx* ctx, const double* ref_pts, int64_t n, double cell, void* stream_)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_nn_index_build: ctx is NULL");
    if (n < 0 || (n > 0 && !ref_pts) || !(cell >= 0.0)) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_nn_index_build: bad arguments");
    if (n >= (int64_t)1 << 31) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_nn_index_build: n must be < 2^31");
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    ctx->nn_ready = false;
    ctx->nn_generation++;
    ctx->nn_n = n;
    ctx->nn_nb[0] = ctx->nn_nb[1] = ctx->nn_nb[2] = 0;
    if (n == 0) { ctx->nn_ready = true; return LRC_OK; }
    const int TB = 256;
    const unsigned gN = (unsigned)((n + TB - 1) / TB);
    int rc = lrc_grow(ctx, &ctx->nn_meta, &ctx->nn_meta_bytes, sizeof(NnMeta));
    if (rc) return rc;
    NnMeta* meta = (NnMeta*)ctx->nn_meta;
    k_nn_meta_init<<<1, 1, 0, stream>>>(meta);
    LRC_CHECK_LAUNCH(ctx, "k_nn_meta_init");
    k_nn_bounds<<<gN, TB, 0, stream>>>(ref_pts, n, meta);
    LRC_CHECK_LAUNCH(ctx, "k_nn_bounds");
    NnMeta hm;
    LRC_CUDA(ctx, cudaMemcpyAsync(&hm, meta, sizeof hm, cudaMemcpyDeviceToHost, stream));
    LRC_CUDA(ctx, cudaStreamSynchronize(stream));
    if (hm.n_finite == 0) { ctx->nn_ready = true; return LRC_OK; }
    double o[3], ext[3];
    for (int k = 0; k < 3; ++k) { o[k] = nn_ord2d(hm.lo[k]); ext[k] = nn_ord2d(hm.hi[k]) - o[k]; }
    double c = cell;
    if (c <= 0.0) {
        // automatic: ~2 points per cell if the points filled the bounding box; surface-like data is sparser, which
        // only means more empty cells
        const double vol = fmax(ext[0], 1e-3) * fmax(ext[1], 1e-3) * fmax(ext[2], 1e-3);
        c = cbrt(vol * 2.0 / (double)hm.n_finite);
        if (!(c > 1e-6)) c = 1e-6;
    }
    auto bins = [&](double cc) { return (floor(ext[0] / cc) + 1.0) * (floor(ext[1] / cc) + 1.0) * (floor(ext[2] / cc) + 1.0); };
    while (bins(c) > 16777216.0) c *= 1.25;                   // at most 2^24 cells
    NnGrid g;
    for (int k = 0; k < 3; ++k) { g.o[k] = o[k]; g.nb[k] = (int)floor(ext[k] / c) + 1; }
    g.cell = c;
    const int64_t nb = (int64_t)g.nb[0] * g.nb[1] * g.nb[2];
    if ((rc = lrc_grow(ctx, &ctx->nn_start, &ctx->nn_start_bytes, sizeof(int) * (size_t)(3 * nb + 2)))) return rc;
    if ((rc = lrc_grow(ctx, &ctx->nn_sorted, &ctx->nn_sorted_bytes, (sizeof(double) * 3 + sizeof(int32_t)) * (size_t)n + 256))) return rc;
    int* start = (int*)ctx->nn_start;
    int* count = start + nb + 1;
    int* cursor = count + nb;
    double* sorted = (double*)ctx->nn_sorted;
    int32_t* sorted_idx = (int32_t*)((char*)ctx->nn_sorted + align_up(sizeof(double) * 3 * (size_t)n, 256));
    LRC_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int) * (size_t)nb, stream));
    k_nn_count<<<gN, TB, 0, stream>>>(ref_pts, n, g, count);
    LRC_CHECK_LAUNCH(ctx, "k_nn_count");
    k_scan_int<<<1, 1024, 0, stream>>>(count, start, cursor, nb);
    LRC_CHECK_LAUNCH(ctx, "k_scan_int");
    k_nn_scatter<<<gN, TB, 0, stream>>>(ref_pts, n, g, cursor, sorted, sorted_idx);
    LRC_CHECK_LAUNCH(ctx, "k_nn_scatter");
    for (int k = 0; k < 3; ++k) { ctx->nn_o[k] = g.o[k]; ctx->nn_nb[k] = g.nb[k]; }
    ctx->nn_cell = c;
    ctx->nn_ready = true;
    return LRC_OK;
}

extern "C" int lrc_nn_query(lrc_ctx* ctx, const float* query_xyz, int64_t M, int32_t* out_index, double* out_distance,
                            const uint32_t* ref_attr_a, uint32_t* out_attr_a, const uint32_t* ref_attr_b, uint32_t* out_attr_b,
                            void* stream_)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_nn_query: ctx is NULL");
    if (!ctx->nn_ready) return lrc_fail(ctx, LRC_ERR_NO_MESH, "lrc_nn_query: call lrc_nn_index_build first");
    if (M < 0 || (M > 0 && (!query_xyz || !out_index))) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_nn_query: bad arguments");
    if (M == 0) return LRC_OK;
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    NnGrid g;
    for (int k = 0; k < 3; ++k) { g.o[k] = ctx->nn_o[k]; g.nb[k] = ctx->nn_nb[k]; }
    g.cell = ctx->nn_cell;
    const double* sorted = (const double*)ctx->nn_sorted;
    const int32_t* sorted_idx = (const int32_t*)((const char*)ctx->nn_sorted + align_up(sizeof(double) * 3 * (size_t)ctx->nn_n, 256));
    k_nn_query<<<(unsigned)((M + 127) / 128), 128, 0, (cudaStream_t)stream_>>>(query_xyz, M, g, (const int*)ctx->nn_start, sorted, sorted_idx,
                                                                               out_index, out_distance, ref_attr_a, out_attr_a, ref_attr_b, out_attr_b);
    LRC_CHECK_LAUNCH(ctx, "k_nn_query");
    return LRC_OK;
}
