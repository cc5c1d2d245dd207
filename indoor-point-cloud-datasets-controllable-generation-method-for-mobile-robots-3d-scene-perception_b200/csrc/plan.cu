// plan.cu -- GPU support for the caller that produces the poses: the coverage-trajectory planner
// (reference trajectory/auto_trajectory_generator.py).
//
// What the reference does, and what costs it time:
//   _is_point_inside_mesh   :220-238  "does ANY mesh vertex lie inside the closed robot cube [p - r, p + r]" -- a numpy
//                                     pass over ALL V vertices for every grid point (:131-146) and every waypoint of
//                                     every candidate (:352-362): O((cells + waypoints) * V)
//   _build_connectivity_graph :245-258  Python double loop over all pairs of free points: O(n^2)
//   _a_star_search          :413-473  open list as a Python set, min() scan per pop: O(n^2)
// Here:
//   lrc_collision_index_build   vertices binned once on a 2-D grid (count -> scan -> scatter); a query reads <= 3 x 3 bins
//   lrc_collision_query         one warp per query point, exact float64 comparisons in the reference's own form
//   lrc_grid_connectivity       free cells ranked in the reference's order (x-major), neighbours found in a (2w+1)^2
//                               window with the reference's distance arithmetic, CSR output with ascending columns
//   lrc_astar                   host-side binary-heap A* over that CSR graph (the reference's search is host code too)
// All verdicts are integers / booleans and are bit-exact against the reference (tests/golden/plan_*.npz).
#include <algorithm>
#include <cmath>
#include <queue>

#include "common.cuh"
#include "scan_util.cuh"

namespace {

struct CiMeta {                 // device-side header of the collision index
    unsigned long long lo[2], hi[2];   // ordered-uint64 encoded xy bounds of the finite vertices
    int n_finite;
};

__device__ __forceinline__ unsigned long long d2ord(double d)
{
    unsigned long long u = (unsigned long long)__double_as_longlong(d);
    return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
static inline double ord2d_host(unsigned long long u)
{
    unsigned long long b = (u & 0x8000000000000000ull) ? (u & 0x7fffffffffffffffull) : ~u;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

__device__ __forceinline__ bool finite3(double x, double y, double z)
{
    return (x - x == 0.0) && (y - y == 0.0) && (z - z == 0.0);
}

__global__ void k_ci_meta_init(CiMeta* m)
{
    m->lo[0] = m->lo[1] = ~0ull;
    m->hi[0] = m->hi[1] = 0ull;
    m->n_finite = 0;
}

__global__ void k_ci_bounds(const double* __restrict__ v, int64_t V, CiMeta* m)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double x = 0, y = 0, z = 0;
    bool ok = false;
    if (i < V) { x = v[3 * i]; y = v[3 * i + 1]; z = v[3 * i + 2]; ok = finite3(x, y, z); }
    unsigned long long lx = ok ? d2ord(x) : ~0ull, hx = ok ? d2ord(x) : 0ull;
    unsigned long long ly = ok ? d2ord(y) : ~0ull, hy = ok ? d2ord(y) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lx = min(lx, __shfl_xor_sync(0xffffffffu, lx, o)); hx = max(hx, __shfl_xor_sync(0xffffffffu, hx, o));
        ly = min(ly, __shfl_xor_sync(0xffffffffu, ly, o)); hy = max(hy, __shfl_xor_sync(0xffffffffu, hy, o));
    }
    const unsigned any = __ballot_sync(0xffffffffu, ok);
    if (lane_id() == 0 && any) {
        atomicMin(&m->lo[0], lx); atomicMax(&m->hi[0], hx);
        atomicMin(&m->lo[1], ly); atomicMax(&m->hi[1], hy);
        atomicAdd(&m->n_finite, __popc(any));
    }
}

struct CiGrid { double ox, oy, cell; int nbx, nby; };

// monotone in its argument: the same function maps vertices and query-cube corners to bins, so a vertex inside a cube
// always lies in a bin of the cube's bin range
__device__ __forceinline__ int bin_of(double x, double o, double cell, int nb)
{
    const double f = floor((x - o) / cell);
    return f < 0.0 ? 0 : (f >= (double)nb ? nb - 1 : (int)f);
}

__global__ void k_ci_count(const double* __restrict__ v, int64_t V, CiGrid g, int* __restrict__ count)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const double x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2];
    if (!finite3(x, y, z)) return;        // NaN / inf compare false in the reference: such a vertex never collides
    atomicAdd(&count[bin_of(y, g.oy, g.cell, g.nby) * g.nbx + bin_of(x, g.ox, g.cell, g.nbx)], 1);
}

__global__ void k_ci_scatter(const double* __restrict__ v, int64_t V, CiGrid g, int* __restrict__ cursor, double* __restrict__ sorted)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const double x = v[3 * i], y = v[3 * i + 1], z = v[3 * i + 2];
    if (!finite3(x, y, z)) return;
    const int pos = atomicAdd(&cursor[bin_of(y, g.oy, g.cell, g.nby) * g.nbx + bin_of(x, g.ox, g.cell, g.nbx)], 1);
    // the order inside a bin depends on the atomics; queries only ask "any", so results do not
    sorted[3 * (int64_t)pos] = x; sorted[3 * (int64_t)pos + 1] = y; sorted[3 * (int64_t)pos + 2] = z;
}

struct Bounds6 { double v[6]; int use; };

// One warp per query point.  Comparisons in the reference's own form:
//   in bounds (:204-217)   x_min <= p - r  and  p + r <= x_max   (all three axes)
//   collision (:220-238)   v >= p - r  and  v <= p + r           (all three axes, any vertex)
constexpr int CQ_THREADS = 128;
__global__ void __launch_bounds__(CQ_THREADS)
k_ci_query(const double* __restrict__ pts, int64_t Q, double half, Bounds6 b, CiGrid g, const int* __restrict__ start,
           const double* __restrict__ sorted, uint8_t* __restrict__ state)
{
    const int64_t q = ((int64_t)blockIdx.x * CQ_THREADS + threadIdx.x) >> 5;
    if (q >= Q) return;
    const int lane = threadIdx.x & 31;
    const double px = pts[3 * q], py = pts[3 * q + 1], pz = pts[3 * q + 2];
    const double lx = __dsub_rn(px, half), ly = __dsub_rn(py, half), lz = __dsub_rn(pz, half);
    const double hx = __dadd_rn(px, half), hy = __dadd_rn(py, half), hz = __dadd_rn(pz, half);
    if (b.use && !(b.v[0] <= lx && hx <= b.v[1] && b.v[2] <= ly && hy <= b.v[3] && b.v[4] <= lz && hz <= b.v[5])) {
        if (lane == 0) state[q] = 0;
        return;
    }
    bool hit = false;
    if (g.nbx > 0 && lx == lx && hx == hx && ly == ly && hy == hy) {
        const int bx0 = bin_of(lx, g.ox, g.cell, g.nbx), bx1 = bin_of(hx, g.ox, g.cell, g.nbx);
        const int by0 = bin_of(ly, g.oy, g.cell, g.nby), by1 = bin_of(hy, g.oy, g.cell, g.nby);
        for (int by = by0; by <= by1 && !hit; ++by) {
            const int s = start[by * g.nbx + bx0], e = start[by * g.nbx + bx1 + 1];   // bins of one row are contiguous
            for (int k0 = s; k0 < e; k0 += 32) {
                const int k = k0 + lane;
                bool in = false;
                if (k < e) {
                    const double vx = sorted[3 * (int64_t)k], vy = sorted[3 * (int64_t)k + 1], vz = sorted[3 * (int64_t)k + 2];
                    in = vx >= lx && vx <= hx && vy >= ly && vy <= hy && vz >= lz && vz <= hz;
                }
                if (__any_sync(0xffffffffu, in)) { hit = true; break; }
            }
        }
    }
    if (lane == 0) state[q] = hit ? 1 : 2;
}

// ---- connectivity graph of the free grid points ----------------------------------------------------------
__global__ void k_cg_flags(const uint8_t* __restrict__ state, int64_t n, int* __restrict__ flag)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = state[i] == 2 ? 1 : 0;
}

__global__ void k_cg_rank(const uint8_t* __restrict__ state, const int* __restrict__ excl, int64_t n, int32_t* __restrict__ free_index)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) free_index[i] = state[i] == 2 ? excl[i] : -1;
}

// ||p_i - p_j|| <= max_dist with the reference's arithmetic (:253): difference first, then sqrt of the sum of squares
// (z is the same constant for every grid point, so its difference is exactly 0)
__device__ __forceinline__ bool cg_near(double xi, double yi, double xj, double yj, double max_dist)
{
    const double dx = __dsub_rn(xi, xj), dy = __dsub_rn(yi, yj);
    const double d = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), 0.0));
    return d <= max_dist;
}

template <bool FILL>
__global__ void k_cg_neighbours(const double* __restrict__ xs, int nx, const double* __restrict__ ys, int ny,
                                const int32_t* __restrict__ free_index, double max_dist, int window,
                                int* __restrict__ count, const int* __restrict__ row_ptr, int32_t* __restrict__ col, int64_t col_cap)
{
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= (int64_t)nx * ny) return;
    const int me = free_index[cell];
    if (me < 0) return;
    const int ix = (int)(cell / ny), iy = (int)(cell - (int64_t)ix * ny);
    const double xi = xs[ix], yi = ys[iy];
    int n = 0;
    int64_t w = FILL ? row_ptr[me] : 0;
    // ascending (ix', iy') == ascending free index == the reference's ascending j
    for (int jx = max(ix - window, 0); jx <= min(ix + window, nx - 1); ++jx)
        for (int jy = max(iy - window, 0); jy <= min(iy + window, ny - 1); ++jy) {
            if (jx == ix && jy == iy) continue;
            const int other = free_index[(int64_t)jx * ny + jy];
            if (other < 0 || !cg_near(xi, yi, xs[jx], ys[jy], max_dist)) continue;
            if (FILL) { if (w < col_cap) col[w] = other; ++w; }
            ++n;
        }
    if (!FILL) count[me] = n;
}

}  // namespace

// ==========================================================================================================
extern "C" int lrc_collision_index_build(lrc_ctx* ctx, const double* verts, int64_t V, double cell, void* stream_)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_collision_index_build: ctx is NULL");
    if (V < 0 || (V > 0 && !verts) || !(cell > 0.0)) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_collision_index_build: bad arguments");
    if (V >= (int64_t)1 << 31) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_collision_index_build: V must be < 2^31");
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    ctx->ci_ready = false;
    ctx->ci_generation++;
    ctx->ci_nbx = ctx->ci_nby = 0;
    ctx->ci_V = V;
    if (V == 0) { ctx->ci_ready = true; return LRC_OK; }
    const int TB = 256;
    const unsigned gV = (unsigned)((V + TB - 1) / TB);
    int rc = lrc_grow(ctx, &ctx->ci_meta, &ctx->ci_meta_bytes, sizeof(CiMeta));
    if (rc) return rc;
    CiMeta* meta = (CiMeta*)ctx->ci_meta;
    k_ci_meta_init<<<1, 1, 0, stream>>>(meta);
    LRC_CHECK_LAUNCH(ctx, "k_ci_meta_init");
    k_ci_bounds<<<gV, TB, 0, stream>>>(verts, V, meta);
    LRC_CHECK_LAUNCH(ctx, "k_ci_bounds");
    CiMeta hm;
    LRC_CUDA(ctx, cudaMemcpyAsync(&hm, meta, sizeof hm, cudaMemcpyDeviceToHost, stream));
    LRC_CUDA(ctx, cudaStreamSynchronize(stream));
    if (hm.n_finite == 0) { ctx->ci_ready = true; return LRC_OK; }
    const double ox = ord2d_host(hm.lo[0]), oy = ord2d_host(hm.lo[1]);
    const double ex = ord2d_host(hm.hi[0]) - ox, ey = ord2d_host(hm.hi[1]) - oy;
    // at most 2^22 bins: a coarser cell only makes a query read more vertices, never fewer
    double c = cell;
    while ((floor(ex / c) + 1.0) * (floor(ey / c) + 1.0) > 4194304.0) c *= 2.0;
    const int nbx = (int)floor(ex / c) + 1, nby = (int)floor(ey / c) + 1;
    const int64_t nb = (int64_t)nbx * nby;
    if ((rc = lrc_grow(ctx, &ctx->ci_start, &ctx->ci_start_bytes, sizeof(int) * (size_t)(3 * nb + 2)))) return rc;
    if ((rc = lrc_grow(ctx, &ctx->ci_sorted, &ctx->ci_sorted_bytes, sizeof(double) * 3 * (size_t)V))) return rc;
    int* start = (int*)ctx->ci_start;            // nb + 1
    int* count = start + nb + 1;                 // nb
    int* cursor = count + nb;                    // nb
    CiGrid g = {ox, oy, c, nbx, nby};
    LRC_CUDA(ctx, cudaMemsetAsync(count, 0, sizeof(int) * (size_t)nb, stream));
    k_ci_count<<<gV, TB, 0, stream>>>(verts, V, g, count);
    LRC_CHECK_LAUNCH(ctx, "k_ci_count");
    k_scan_int<<<1, 1024, 0, stream>>>(count, start, cursor, nb);
    LRC_CHECK_LAUNCH(ctx, "k_scan_int");
    k_ci_scatter<<<gV, TB, 0, stream>>>(verts, V, g, cursor, (double*)ctx->ci_sorted);
    LRC_CHECK_LAUNCH(ctx, "k_ci_scatter");
    ctx->ci_ox = ox; ctx->ci_oy = oy; ctx->ci_cell = c; ctx->ci_nbx = nbx; ctx->ci_nby = nby;
    ctx->ci_ready = true;
    return LRC_OK;
}

extern "C" int lrc_collision_query(lrc_ctx* ctx, const double* pts, int64_t Q, double half, const double* h_bounds,
                                   uint8_t* state, void* stream_)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_collision_query: ctx is NULL");
    if (!ctx->ci_ready) return lrc_fail(ctx, LRC_ERR_NO_MESH, "lrc_collision_query: call lrc_collision_index_build first");
    if (Q < 0 || (Q > 0 && (!pts || !state)) || !(half >= 0.0)) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_collision_query: bad arguments");
    if (Q == 0) return LRC_OK;
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    Bounds6 b;
    b.use = h_bounds != nullptr;
    for (int k = 0; k < 6; ++k) b.v[k] = h_bounds ? h_bounds[k] : 0.0;
    CiGrid g = {ctx->ci_ox, ctx->ci_oy, ctx->ci_cell, ctx->ci_nbx, ctx->ci_nby};
    const int64_t threads = Q * 32;
    k_ci_query<<<(unsigned)((threads + CQ_THREADS - 1) / CQ_THREADS), CQ_THREADS, 0, (cudaStream_t)stream_>>>(
        pts, Q, half, b, g, (const int*)ctx->ci_start, (const double*)ctx->ci_sorted, state);
    LRC_CHECK_LAUNCH(ctx, "k_ci_query");
    return LRC_OK;
}

extern "C" int lrc_grid_connectivity(lrc_ctx* ctx, const double* xs, int32_t nx, const double* ys, int32_t ny,
                                     const uint8_t* state, double max_dist, int32_t window, int32_t* free_index,
                                     int32_t* row_ptr, int32_t* col, int64_t col_capacity, int64_t* h_counts, void* stream_)
{
    if (!ctx) return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_grid_connectivity: ctx is NULL");
    if (nx < 0 || ny < 0 || window < 0 || !h_counts || !(max_dist >= 0.0))
        return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_grid_connectivity: bad arguments");
    const int64_t n = (int64_t)nx * ny;
    h_counts[0] = h_counts[1] = 0;
    if (n == 0) return LRC_OK;
    if (!xs || !ys || !state || !free_index || !row_ptr) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_grid_connectivity: NULL array");
    if (n >= (int64_t)1 << 30) return lrc_fail(ctx, LRC_ERR_INVALID, "lrc_grid_connectivity: grid too large");
    LRC_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t stream = (cudaStream_t)stream_;
    int rc = lrc_grow(ctx, &ctx->cg_scratch, &ctx->cg_scratch_bytes, sizeof(int) * (size_t)(3 * n + 4));
    if (rc) return rc;
    int* flag = (int*)ctx->cg_scratch;        // n
    int* excl = flag + n;                     // n + 1
    int* count = excl + n + 1;                // n (indexed by free rank)
    const int TB = 256;
    const unsigned gN = (unsigned)((n + TB - 1) / TB);
    k_cg_flags<<<gN, TB, 0, stream>>>(state, n, flag);
    LRC_CHECK_LAUNCH(ctx, "k_cg_flags");
    k_scan_int<<<1, 1024, 0, stream>>>(flag, excl, nullptr, n);
    LRC_CHECK_LAUNCH(ctx, "k_scan_int");
    k_cg_rank<<<gN, TB, 0, stream>>>(state, excl, n, free_index);
    LRC_CHECK_LAUNCH(ctx, "k_cg_rank");
    int n_free = 0;
    LRC_CUDA(ctx, cudaMemcpyAsync(&n_free, excl + n, sizeof(int), cudaMemcpyDeviceToHost, stream));
    LRC_CUDA(ctx, cudaStreamSynchronize(stream));
    h_counts[0] = n_free;
    if (n_free == 0) { LRC_CUDA(ctx, cudaMemsetAsync(row_ptr, 0, sizeof(int32_t), stream)); return LRC_OK; }
    k_cg_neighbours<false><<<gN, TB, 0, stream>>>(xs, nx, ys, ny, free_index, max_dist, window, count, nullptr, nullptr, 0);
    LRC_CHECK_LAUNCH(ctx, "k_cg_neighbours");
    k_scan_int<<<1, 1024, 0, stream>>>(count, row_ptr, nullptr, n_free);
    LRC_CHECK_LAUNCH(ctx, "k_scan_int");
    int nnz = 0;
    LRC_CUDA(ctx, cudaMemcpyAsync(&nnz, row_ptr + n_free, sizeof(int), cudaMemcpyDeviceToHost, stream));
    LRC_CUDA(ctx, cudaStreamSynchronize(stream));
    h_counts[1] = nnz;
    if (nnz > col_capacity || (nnz > 0 && !col)) return lrc_fail(ctx, LRC_ERR_CAPACITY, "lrc_grid_connectivity: col_capacity is smaller than the number of edges");
    if (nnz > 0) {
        k_cg_neighbours<true><<<gN, TB, 0, stream>>>(xs, nx, ys, ny, free_index, max_dist, window, nullptr, row_ptr, col, col_capacity);
        LRC_CHECK_LAUNCH(ctx, "k_cg_neighbours");
    }
    return LRC_OK;
}

// A* over a CSR graph of points (host).  Edge cost and heuristic are the Euclidean distance, as in the reference
// (:430-434,:457); closed nodes are never reopened (:452-453).  Ties on f are broken by the smaller node index, so the
// result is deterministic; among equal-cost paths the reference's choice depends on CPython's set iteration order.
extern "C" int lrc_astar(const int32_t* h_row_ptr, const int32_t* h_col, const double* h_pts, int32_t n, int32_t start,
                         int32_t end, int32_t* h_path, int32_t path_capacity, int32_t* h_path_len, double* h_cost)
{
    if (!h_row_ptr || !h_pts || !h_path_len || n <= 0 || start < 0 || start >= n || end < 0 || end >= n)
        return lrc_fail(nullptr, LRC_ERR_INVALID, "lrc_astar: bad arguments");
    *h_path_len = 0;
    if (h_cost) *h_cost = 0.0;
    auto dist = [&](int a, int b) {
        const double dx = h_pts[3 * a] - h_pts[3 * b], dy = h_pts[3 * a + 1] - h_pts[3 * b + 1], dz = h_pts[3 * a + 2] - h_pts[3 * b + 2];
        return std::sqrt(dx * dx + dy * dy + dz * dz);
    };
    if (start == end) {
        if (path_capacity < 1 || !h_path) return lrc_fail(nullptr, LRC_ERR_CAPACITY, "lrc_astar: path_capacity too small");
        h_path[0] = start;
        *h_path_len = 1;
        return LRC_OK;
    }
    const double INF = INFINITY;
    std::vector<double> g((size_t)n, INF);
    std::vector<int32_t> came((size_t)n, -1);
    std::vector<unsigned char> closed((size_t)n, 0);
    typedef std::pair<double, int32_t> Item;                       // (f, node): smaller f first, then smaller index
    std::priority_queue<Item, std::vector<Item>, std::greater<Item>> open;
    g[start] = 0.0;
    open.push(Item(dist(start, end), start));
    bool found = false;
    while (!open.empty()) {
        const Item it = open.top();
        open.pop();
        const int32_t cur = it.second;
        if (closed[cur]) continue;                                 // stale entry
        if (cur == end) { found = true; break; }
        closed[cur] = 1;
        for (int32_t e = h_row_ptr[cur]; e < h_row_ptr[cur + 1]; ++e) {
            const int32_t nb = h_col[e];
            if (nb < 0 || nb >= n || closed[nb]) continue;
            const double tg = g[cur] + dist(cur, nb);
            if (tg >= g[nb]) continue;
            g[nb] = tg;
            came[nb] = cur;
            open.push(Item(tg + dist(nb, end), nb));
        }
    }
    if (!found) return LRC_OK;                                     // *h_path_len == 0: no path (:473)
    int32_t len = 0;
    for (int32_t c = end; c >= 0; c = came[c]) ++len;
    if (len > path_capacity || !h_path) return lrc_fail(nullptr, LRC_ERR_CAPACITY, "lrc_astar: path_capacity too small");
    int32_t k = len;
    for (int32_t c = end; c >= 0; c = came[c]) h_path[--k] = c;
    *h_path_len = len;
    if (h_cost) *h_cost = g[end];
    return LRC_OK;
}
